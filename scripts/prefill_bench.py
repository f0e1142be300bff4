#!/usr/bin/env python
"""Prefill microbenchmark (BASELINE config 5 shape per sequence: prompt 4096, LLaMA-2-7B heads): the tensor-core
prefill attention kernel on SEQS fresh prompts of LEN tokens, and the W8A8 GEMMs at M = SEQS * LEN tokens through
scripts/gemm_bench.py.  Prints attention ms and TFLOP/s (causal FLOPs 2 * n^2 * D per head for QK^T + PV)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200_import  # noqa: E402

b200_import.load()
from ppl_llm_serving_b200 import capi  # noqa: E402
from ppl_llm_serving_b200.engine import _ptr  # noqa: E402

lib = capi.load_library()
SEQS, LEN, H, D = int(os.environ.get("SEQS", 8)), int(os.environ.get("LEN", 4096)), 32, 128
IMPL = int(os.environ.get("IMPL", 2))  # 2: mma.sync prefill kernel, 6: tcgen05 / TMEM kernel (experimental)
T = SEQS * LEN
geom = capi.KvGeomC()
geom.num_layers, geom.num_kv_heads, geom.head_dim, geom.quant_group = 1, H, D, 8
geom.cache_layout, geom.cache_mode, geom.page_size, geom.max_tokens = 3, 0, 16, T
cache = torch.zeros((2 * H * T * D,), dtype=torch.int8, device="cuda")
scale = torch.zeros((2 * H * T * D // 8,), dtype=torch.float16, device="cuda")
qkv = torch.randn((T, 3 * H * D), dtype=torch.float16, device="cuda")
out = torch.empty((T, H * D), dtype=torch.float16, device="cuda")
seq_starts = (torch.arange(SEQS + 1, dtype=torch.int64) * LEN).cuda()
start_pos = torch.zeros(SEQS, dtype=torch.int64, device="cuda")
idx = (torch.arange(SEQS, dtype=torch.int64) * LEN).cuda()
tok = torch.zeros(T, dtype=torch.int64, device="cuda")
st = capi.StepC()
st.token_ids, st.seq_starts, st.kv_starts = tok.data_ptr(), seq_starts.data_ptr(), seq_starts.data_ptr()
st.cache_indices, st.start_pos = idx.data_ptr(), start_pos.data_ptr()
st.num_tokens, st.batch, st.decoding_batches = T, SEQS, 0
st.max_seq_len, st.max_kv_len, st.max_pages = LEN, LEN, 0
ws = torch.empty(lib.b2llm_attention_workspace_size(SEQS, H, D), dtype=torch.uint8, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run():
    rc = lib.b2llm_op_attention(sp, _ptr(qkv), C.byref(st), H, C.byref(geom), 0, _ptr(cache), _ptr(scale), _ptr(ws), _ptr(out), IMPL)
    assert rc == 0, lib.b2llm_last_error()


for _ in range(2):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 5
e0.record()
for _ in range(reps):
    run()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
flops = 2.0 * LEN * LEN * D * H * SEQS
print(f"prefill attention: {SEQS} x {LEN} tokens, {H} heads: {ms:.3f} ms, {flops / ms / 1e9:.1f} TFLOP/s (causal), "
      f"{T / ms * 1e3:.0f} tokens/s/layer")
