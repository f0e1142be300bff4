#!/usr/bin/env python
"""Per-CTA timeline of ONE decode-attention launch (b2llm_debug_attention_trace): where do the ~27 us a launch costs beyond
bytes / bandwidth go -- ramp, tail, or the CTAs' own prologue?  Same shapes / env as attn_bench.py (B, H, HKV, KV)."""
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
B, KV, D, PAGE = int(os.environ.get("B", 1024)), int(os.environ.get("KV", 512)), 128, 16
HQ = int(os.environ.get("H", 32))
H = int(os.environ.get("HKV", HQ))
T = B * KV
L = max(2, min(8, int(12e9 // (2 * H * T * D * 1.25))))
geom = capi.KvGeomC()
geom.num_layers, geom.num_kv_heads, geom.head_dim, geom.quant_group = L, H, D, 8
geom.cache_layout, geom.cache_mode, geom.page_size, geom.max_tokens = 3, 1, PAGE, T
cache = torch.randint(-127, 128, (L * 2 * H * T * D,), dtype=torch.int8, device="cuda")
scale = torch.full((L * 2 * H * T * D // 8,), 0.01, dtype=torch.float16, device="cuda")
qkv = torch.randn((B, (HQ + 2 * H) * D), dtype=torch.float16, device="cuda")
out = torch.empty((B, HQ * D), dtype=torch.float16, device="cuda")
pages_per = KV // PAGE
perm = np.random.default_rng(0).permutation(B * pages_per)
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
    rc = lib.b2llm_op_attention(sp, _ptr(qkv), C.byref(st), HQ, C.byref(geom), layer % L, _ptr(cache), _ptr(scale), _ptr(ws), _ptr(out), 2)
    assert rc == 0, lib.b2llm_last_error()


CAP = 1 << 18
trace = torch.zeros(4 * CAP, dtype=torch.int64, device="cuda")
for l in range(4):
    run(l)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
lib.b2llm_debug_attention_trace(_ptr(trace), CAP)
e0.record()
run(5)
e1.record()
torch.cuda.synchronize()
lib.b2llm_debug_attention_trace(None, 0)
tr = trace.cpu().numpy().reshape(-1, 4)
tr = tr[tr[:, 2] != 0]
t0 = tr[:, 0].min()
start, loop, end, sm = (tr[:, 0] - t0) * 1e-3, (tr[:, 1] - t0) * 1e-3, (tr[:, 2] - t0) * 1e-3, tr[:, 3]
total = end.max()
ns, nw = C.c_int32(), C.c_int32()
lib.b2llm_attention_decode_plan(B, HQ, H, KV, C.byref(ns), C.byref(nw))
slots = 148 * 12 // nw.value
bytes_alg = B * KV * 2 * H * D * 1.25
print(f"B={B} H={HQ}/{H} KV={KV}: plan {ns.value} splits x {nw.value} warps, {len(tr)} CTAs on {len(np.unique(sm))} SMs, {slots} CTA slots; "
      f"event time (kernel + merge) {e0.elapsed_time(e1) * 1e3:.1f} us, first CTA start -> last CTA end {total:.1f} us "
      f"= {bytes_alg / total / 1e3:.0f} GB/s; bytes / 7.08 TB/s = {bytes_alg / 7.08e6:.1f} us")
q = lambda a: " ".join(f"{np.percentile(a, p):.1f}" for p in (0, 10, 50, 90, 100))
print(f"  CTA start (us, pct 0/10/50/90/100): {q(start)};  prologue start -> main loop: {q(loop - start)};  lifetime: {q(end - start)}")
first_wave = np.sort(start)[:slots]
print(f"  first {slots} CTAs all started by {first_wave.max():.1f} us; first CTA ended at {end.min():.1f} us; last CTA STARTED at {start.max():.1f} us")
print(f"  slot utilisation sum(lifetime) / (slots x span) = {(end - start).sum() / (slots * total):.3f}; "
      f"in-loop share of lifetime = {(end - loop).sum() / (end - start).sum():.3f}")
edges = np.arange(0, total + 5, 5.0)
active = [int(((start < b) & (end > a)).sum()) for a, b in zip(edges[:-1], edges[1:])]
print("  active CTAs per 5 us bin: " + " ".join(str(a) for a in active))
per_sm_last = np.array([end[sm == s].max() for s in np.unique(sm)])
print(f"  per-SM time of last CTA end (us): {q(per_sm_last)}")
