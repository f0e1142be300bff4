"""Import helper: the package directory is literally ``ppl.llm.serving_b200/`` (dots), which Python's
import system cannot address by name; register it as ``ppl_llm_serving_b200``."""
import importlib.util
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent
PKG_DIR = ROOT / "ppl.llm.serving_b200"
NAME = "ppl_llm_serving_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, PKG_DIR / "__init__.py",
                                                  submodule_search_locations=[str(PKG_DIR)])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod
