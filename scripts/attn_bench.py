#!/usr/bin/env python
"""Decode-attention microbenchmark at the bench shape (B=1024, kv_len 512, 7B MHA, int8 group-8 paged KV):
the attention op alone in a loop over L distinct layers of a real-size cache (so every launch streams fresh HBM),
CUDA events per launch.  Prints ms/launch and achieved algorithmic GB/s."""
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
# B, H (q heads), HKV (kv heads), KV: the per-rank shapes of the tensor-parallel configs can be timed on one GPU, e.g.
#   7B TP 8: H=4 HKV=4;  70B TP 8: B=256 H=8 HKV=1 KV=8192;  13B TP 4: B=512 H=10 HKV=10 KV=2640
# B2LLM_ATTN_SPLITS / B2LLM_ATTN_WARPS override the kernel's launch plan (plan_decode) for experiments.
B, KV, D, PAGE = int(os.environ.get("B", 1024)), int(os.environ.get("KV", 512)), 128, 16
HQ = int(os.environ.get("H", 32))
H = int(os.environ.get("HKV", HQ))
T = B * KV
L = max(1, min(32, int(12e9 // (2 * H * T * D * 1.25))))   # layers of cache to cycle through (<= 12 GB)
geom = capi.KvGeomC()
geom.num_layers, geom.num_kv_heads, geom.head_dim, geom.quant_group = L, H, D, 8
geom.cache_layout, geom.cache_mode, geom.page_size, geom.max_tokens = 3, 1, PAGE, T
cache = torch.randint(-127, 128, (L * 2 * H * T * D,), dtype=torch.int8, device="cuda")
scale = torch.full((L * 2 * H * T * D // 8,), 0.01, dtype=torch.float16, device="cuda")
qkv = torch.randn((B, (HQ + 2 * H) * D), dtype=torch.float16, device="cuda")
out = torch.empty((B, HQ * D), dtype=torch.float16, device="cuda")
rng = np.random.default_rng(0)
pages_per = KV // PAGE
perm = rng.permutation(B * pages_per)
page_list = torch.from_numpy((perm.reshape(B, pages_per) * PAGE).astype(np.int64)).cuda()
seq_starts = torch.arange(B + 1, dtype=torch.int64, device="cuda")
start_pos = torch.full((B,), KV - 1, dtype=torch.int64, device="cuda")
kv_starts = seq_starts * KV
tok = torch.zeros(B, dtype=torch.int64, device="cuda")
st = capi.StepC()
st.token_ids, st.seq_starts, st.kv_starts = tok.data_ptr(), seq_starts.data_ptr(), kv_starts.data_ptr()
st.cache_indices, st.start_pos = page_list.data_ptr(), start_pos.data_ptr()
st.num_tokens, st.batch, st.decoding_batches = B, B, B
st.max_seq_len, st.max_kv_len, st.max_pages = 1, KV, pages_per
ws = torch.empty(lib.b2llm_attention_workspace_size(B, HQ, D), dtype=torch.uint8, device="cuda")
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)


def run(layer):
    rc = lib.b2llm_op_attention(sp, _ptr(qkv), C.byref(st), HQ, C.byref(geom), layer, _ptr(cache), _ptr(scale), _ptr(ws), _ptr(out), 2)
    assert rc == 0, lib.b2llm_last_error()


for l in range(L):
    run(l)
torch.cuda.synchronize()
reps = 3
evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps * L)]
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for i in range(reps * L):
    evs[i][0].record()
    run(i % L)
    evs[i][1].record()
t1.record()
torch.cuda.synchronize()
ms = np.array([a.elapsed_time(b) for a, b in evs])
bytes_per = B * KV * 2 * H * D * 1.25
ns, nw = C.c_int32(), C.c_int32()
lib.b2llm_attention_decode_plan(B, HQ, H, KV, C.byref(ns), C.byref(nw))
print(f"B={B} H={HQ}/{H} KV={KV} plan(default)=({ns.value} splits, {nw.value} warps) "
      f"override=({os.environ.get('B2LLM_ATTN_SPLITS', '-')}, {os.environ.get('B2LLM_ATTN_WARPS', '-')}) ", end="")
print(f"attention alone: per-launch median {np.median(ms):.4f} ms min {ms.min():.4f} max {ms.max():.4f}; "
      f"loop avg {t0.elapsed_time(t1) / (reps * L):.4f} ms -> {bytes_per / (t0.elapsed_time(t1) / (reps * L) * 1e-3) / 1e9:.0f} GB/s "
      f"(median {bytes_per / (np.median(ms) * 1e-3) / 1e9:.0f} GB/s)")
