#!/usr/bin/env python
"""BASELINE config 5 ("Prefill microbench: batch 64 x prompt 4096, LLaMA-7B W8A8, TP=1 -- prefill tokens/sec and GEMM
tensor-pipe %"): ONE whole prefill step of SEQS fresh prompts of LEN tokens through the engine (32 layers, embedding ->
logits of the last token of every prompt), device-timed, with the per-class split the engine's profiler gives
(attention = prefill flash-attention kernel; layer GEMMs = W8A8 tcgen05 kernels).

    SEQS=64 LEN=4096 python scripts/prefill_step_bench.py     # the literal config-5 shape (262 144 tokens / step)

This is a kernel-level run, not a scheduler run: the reference's generator caps a step at --max-tokens-per-step
(default 8192, tools/offline_inference.cc:56).  One JSON line on stdout.
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
from ppl_llm_serving_b200.engine import CudaResourceManager, LLMEngine, ModelConfig, ModelInput, LLAMA2_7B, RC_SUCCESS  # noqa: E402

SEQS, LEN = int(os.environ.get("SEQS", 16)), int(os.environ.get("LEN", 4096))
LAYERS, STEPS, PAGE = int(os.environ.get("LAYERS", 32)), int(os.environ.get("STEPS", 2)), 16
T = SEQS * LEN

cfg = ModelConfig(**LLAMA2_7B, page_size=PAGE, max_position=max(4096, LEN))
cfg.num_layers = LAYERS
res = CudaResourceManager()
rc = res.Init(cfg, 0.9, max_running_batch=SEQS, max_tokens_per_step=T, enable_penalty=False, kv_cache_max_tokens=T,
              seed=0xB200, device=0)
assert rc == RC_SUCCESS, res.lib.b2llm_last_error()
lib = res.lib
engine = LLMEngine(res, False, 1, 0.0)
rng = np.random.default_rng(1005)
pages_per = LEN // PAGE
mi = ModelInput()
mi.token_inputs = rng.integers(0, cfg.vocab_size, T).astype(np.int64)
mi.seq_starts = np.arange(SEQS + 1, dtype=np.int64) * LEN
mi.kv_starts = np.arange(SEQS + 1, dtype=np.int64) * LEN
mi.start_pos = np.zeros(SEQS, dtype=np.int64)
mi.page_list = (rng.permutation(SEQS * pages_per).reshape(SEQS, pages_per) * PAGE).astype(np.int64).reshape(-1)
mi.max_pages, mi.decoding_batches, mi.max_seq_len, mi.max_kv_len = pages_per, 0, LEN, LEN
assert engine.SetInput(mi, True) == RC_SUCCESS, lib.b2llm_last_error()

stream = res.stream


def step():
    rc = engine.RunModel(False)
    assert rc == RC_SUCCESS, lib.b2llm_last_error()


step()  # warm-up (also sizes the activation buffers)
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
attn_ms, gemm_ms = ms_cls[0] / STEPS, ms_cls[1] / STEPS
h, I = cfg.hidden_dim, cfg.intermediate_dim
gemm_flops = 2.0 * T * LAYERS * (3 * h * h + h * h + 3 * h * I)          # int8 MACs x 2, the four projections
attn_flops = 2.0 * LEN * LEN * cfg.head_dim * cfg.num_heads * SEQS * LAYERS  # causal: QK^T + PV over the lower triangle
peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
bf16 = float(peaks.get("bf16_tflops_sustained", 0) or 0)
print(json.dumps({
    "metric": "prefill tokens/sec, LLaMA-2-7B W8A8 TP=1 (BASELINE config 5)", "value": T / ms * 1e3, "unit": "tokens/s",
    "config": {"workload": f"{SEQS} prompts x {LEN} tokens = {T} tokens per step, {LAYERS} layers, int8 group-8 paged KV "
                           f"(page {PAGE}, layout 3) written by the step, logits of the last token of every prompt"},
    "ms_per_step": ms, "steps": STEPS,
    "device_ms_by_class_per_step": {"attention": attn_ms, "layer_gemms": gemm_ms, "lm_head": ms_cls[2] / STEPS,
                                    "other (norm, quant, rope + KV append, embedding)": ms - attn_ms - gemm_ms - ms_cls[2] / STEPS},
    "gemm": {"int8_pflops": gemm_flops / gemm_ms / 1e12 if gemm_ms else None,
             "frac_of_2x_measured_bf16_sustained": gemm_flops / gemm_ms / 1e9 / (2 * bf16) if gemm_ms and bf16 else None},
    "attention": {"causal_tflops_fp16": attn_flops / attn_ms / 1e9 if attn_ms else None, "kernel": "attn_prefill_kernel (mma.sync)"},
}), flush=True)
res.close()
