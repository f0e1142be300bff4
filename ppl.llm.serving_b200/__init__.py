"""B200-native (sm_100a) implementation of ppl.llm.serving's batched LLaMA decode hot path.

Layout:
  csrc/      hand-written CUDA kernels + the C ABI (include/b2llm.h) -> lib/libb2llm.so
  capi.py    ctypes binding of the C ABI
  engine.py  Python mirror of the reference's engine objects (ModelInput, LLMEngine.Execute,
             CudaPostProcessor, CudaResourceManager) used by the tests and bench.py
  host/      C++ mirror of the same objects (ppl::llm::*), for linking the reference's tools

The directory name contains dots, so it is imported through ``b200_import.py`` at the repo root
under the module name ``ppl_llm_serving_b200``.
"""
from . import capi  # noqa: F401
