#!/usr/bin/env python
"""Decode step of LLaMA-2-7B W8A8 at an arbitrary (running batch, uniform kv_len) -- the shapes of SURVEY 8(d) that
bench.py (fixed at the headline B = 1024) does not cover, e.g. config 2b "kv_len 2048 at the batch that fits":

    BATCH=250 KV_LEN=2048 python scripts/decode_shape_bench.py
    BATCH=64  KV_LEN=8192 PAGE=128 python scripts/decode_shape_bench.py

Device-timed (CUDA events on the engine stream, inputs resident in HBM), per-class split from the engine's profiler,
fraction of the HBM roofline for the step's algorithmic bytes.  One JSON line.
"""
import ctypes as C
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200_import  # noqa: E402

b200_import.load()
from ppl_llm_serving_b200.engine import CudaResourceManager, LLMEngine, ModelConfig, ModelInput, LLAMA2_7B, RC_SUCCESS, _ptr  # noqa: E402

BATCH, KV_LEN = int(os.environ.get("BATCH", 250)), int(os.environ.get("KV_LEN", 2048))
PAGE, LAYERS = int(os.environ.get("PAGE", 16)), int(os.environ.get("LAYERS", 32))
STEPS, WARMUP = int(os.environ.get("STEPS", 10)), int(os.environ.get("WARMUP", 3))
assert KV_LEN % PAGE == 0

cfg = ModelConfig(**LLAMA2_7B, page_size=PAGE, max_position=max(4096, KV_LEN + 1))
cfg.num_layers = LAYERS
res = CudaResourceManager()
rc = res.Init(cfg, 0.94, max_running_batch=BATCH, max_tokens_per_step=BATCH, enable_penalty=False, seed=0xB200, device=0)
assert rc == RC_SUCCESS, res.lib.b2llm_last_error()
lib = res.lib
pages_per = KV_LEN // PAGE
assert BATCH * KV_LEN <= res.kv_cache_max_tokens, f"needs {BATCH * KV_LEN} KV tokens, budget {res.kv_cache_max_tokens}"
res.kv_cache_mem.random_(-127, 128)
res.kv_scale_mem.fill_(0.01)
rng = np.random.default_rng(1003)
engine = LLMEngine(res, False, 1, 0.0)
mi = ModelInput()
mi.token_inputs = rng.integers(0, cfg.vocab_size, BATCH).astype(np.int64)
mi.seq_starts = np.arange(BATCH + 1, dtype=np.int64)
mi.start_pos = np.full(BATCH, KV_LEN - 1, dtype=np.int64)
mi.kv_starts = np.arange(BATCH + 1, dtype=np.int64) * KV_LEN
mi.page_list = (rng.permutation(BATCH * pages_per).reshape(BATCH, pages_per) * PAGE).astype(np.int64).reshape(-1)
mi.max_pages, mi.decoding_batches, mi.max_seq_len, mi.max_kv_len = pages_per, BATCH, 1, KV_LEN
assert engine.SetInput(mi, True) == RC_SUCCESS, lib.b2llm_last_error()
stream = res.stream
sptr = C.c_void_p(stream.cuda_stream)
dev_tok = torch.empty(BATCH, dtype=torch.int32, device="cuda")
dev_lp = torch.empty(BATCH, dtype=torch.float32, device="cuda")


def step():
    assert engine.RunModel(False) == RC_SUCCESS, lib.b2llm_last_error()
    rc = lib.b2llm_sample_topk_topp(sptr, C.c_void_p(engine.logits_ptr), None, None, None, BATCH, cfg.vocab_size,
                                    engine.logits_stride, 1, 0.0, 0.0, None, _ptr(dev_tok), _ptr(dev_lp))
    assert rc == RC_SUCCESS


for _ in range(WARMUP):
    step()
torch.cuda.synchronize()
lib.b2llm_engine_profile(res.engine, 1)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record(stream)
for _ in range(STEPS):
    step()
ev1.record(stream)
torch.cuda.synchronize()
ms = ev0.elapsed_time(ev1) / STEPS
ms_cls = (C.c_double * 3)()
n_cls = (C.c_int64 * 3)()
lib.b2llm_engine_profile_read(res.engine, ms_cls, n_cls, 3)
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
hbm = float(peaks.get("hbm_gbs", 6650.0))
h, I = cfg.hidden_dim, cfg.intermediate_dim
attn_bytes = BATCH * KV_LEN * 2 * cfg.num_kv_heads * cfg.head_dim * (1 + 2 / cfg.cache_quant_group)
step_bytes = LAYERS * (4 * h * h + 3 * h * I + attn_bytes) + cfg.vocab_size * h * 2
attn_ms = ms_cls[0] / max(1, n_cls[0])
print(json.dumps({
    "metric": "decode tokens/sec, LLaMA-2-7B W8A8 TP=1", "value": BATCH / ms * 1e3, "unit": "tokens/s", "ms_per_step": ms,
    "config": {"workload": f"running batch {BATCH}, uniform kv_len {KV_LEN}, {LAYERS} layers, int8 group-8 paged KV page_size {PAGE} layout 3, greedy",
               "kv_budget_tokens": res.kv_cache_max_tokens},
    "device_ms_by_class_per_step": {"attention": ms_cls[0] / STEPS, "layer_gemms": ms_cls[1] / STEPS, "lm_head": ms_cls[2] / STEPS},
    "step_roofline": {"algorithmic_bytes_per_step": step_bytes, "hbm_bound_ms": step_bytes / (hbm * 1e9) * 1e3,
                      "frac_of_hbm_roofline": step_bytes / (hbm * 1e9) * 1e3 / ms},
    "attention": {"avg_launch_ms": attn_ms, "GB/s": attn_bytes / (attn_ms * 1e-3) / 1e9 if attn_ms else None,
                  "frac_of_hbm_peak": attn_bytes / (attn_ms * 1e-3) / 1e9 / hbm if attn_ms else None},
}), flush=True)
res.close()
