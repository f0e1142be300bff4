#!/usr/bin/env python
"""Scheduler steady state ("2c"): the reference's OWN LLMGenerator / LLMEngine / CudaResourceManager
(oracle/_ref/token_inout_driver, built from the reference's sources in place) over libpplnn_b200.so + libb2llm.so,
LLaMA-2-7B W8A8 random-init, --max-running-batch 1024, synthetic token-in/out requests whose total length makes the
KV budget (0.94 x free memory, the reference's formula) the binding limit.  Prints the driver's [RESULT] line and the
reference's [PERF] profiler dump; one summary JSON line at the end."""
import json
import os
import subprocess
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200_import  # noqa: E402

b200_import.load()
from ppl_llm_serving_b200.engine import ModelConfig, LLAMA2_7B  # noqa: E402
from ppl_llm_serving_b200.model_slice import write_model_dir  # noqa: E402

requests = int(os.environ.get("REQUESTS", 2048))
prompt = int(os.environ.get("PROMPT", 256))
gen = int(os.environ.get("GEN", 256))
layers = int(os.environ.get("LAYERS", 32))
cfg = ModelConfig(**LLAMA2_7B, page_size=16, max_position=4096)
cfg.num_layers = layers
with tempfile.TemporaryDirectory() as td:
    mdir = write_model_dir(Path(td) / "model", cfg, seed=0xB200)
    cmd = [str(ROOT / "oracle" / "_ref" / "token_inout_driver"), "--model-dir", str(mdir), "--quant-method", "online_i8i8",
           "--max-running-batch", "1024", "--max-tokens-per-step", "8192", "--max-tokens-scale", "0.94",
           "--max-total-tokens-per-request", "8192", "--requests", str(requests), "--prompt-len", str(prompt),
           "--gen-len", str(gen), "--enable-profiling", "1"]
    env = dict(os.environ, PPL_LOG_LEVEL="WARNING")
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=1500)
perf = [l for l in r.stderr.splitlines() if "[PERF]" in l or "failed" in l.lower() or "error" in l.lower()]
print("\n".join(perf[-60:]))
res = [l for l in r.stdout.splitlines() if l.startswith("[RESULT]")]
print(r.stdout[-500:] if not res else res[-1])
if res:
    d = json.loads(res[-1].split("[RESULT]")[1])
    d.update(workload=f"7B W8A8 {layers} layers, {requests} requests x (prompt {prompt} + gen {gen}), max-running-batch 1024, "
                      f"through the reference's LLMGenerator (C++ host path)", rc=r.returncode)
    print(json.dumps(d))
