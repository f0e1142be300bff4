#!/usr/bin/env python
"""Feasibility probe: can decode attention (HBM-bound) and the layer's W8A8 GEMMs (tensor-bound) of TWO half batches
co-run on two streams?  Times, at the bench shape (7B, kv_len 512):
  serial   : attention(B) + GEMM chain(M = B)                                  -- what the engine does today
  halves   : attention(B/2) alone, GEMM chain(M = B/2) alone
  co-run   : attention(B/2) on stream 0  ||  GEMM chain(M = B/2) on stream 1   -- per priority assignment
A two-micro-batch engine would spend 2 x co-run per layer instead of serial.  Nothing here is a bench number."""
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
BFULL, KV, D, PAGE, HQ = int(os.environ.get("B", 1024)), int(os.environ.get("KV", 512)), 128, 16, 32
HID, INTER = 4096, 11008
L = 6                                                   # layers of cache / weight sets cycled through


class Attn:
    def __init__(self, B):
        T = B * KV
        self.B = B
        g = capi.KvGeomC()
        g.num_layers, g.num_kv_heads, g.head_dim, g.quant_group = L, HQ, D, 8
        g.cache_layout, g.cache_mode, g.page_size, g.max_tokens = 3, 1, PAGE, T
        self.geom = g
        self.cache = torch.randint(-127, 128, (L * 2 * HQ * T * D,), dtype=torch.int8, device="cuda")
        self.scale = torch.full((L * 2 * HQ * T * D // 8,), 0.01, dtype=torch.float16, device="cuda")
        self.qkv = torch.randn((B, 3 * HQ * D), dtype=torch.float16, device="cuda")
        self.out = torch.empty((B, HQ * D), dtype=torch.float16, device="cuda")
        pp = KV // PAGE
        perm = np.random.default_rng(0).permutation(B * pp)
        self.page_list = torch.from_numpy((perm.reshape(B, pp) * PAGE).astype(np.int64)).cuda()
        self.seq_starts = torch.arange(B + 1, dtype=torch.int64, device="cuda")
        self.start_pos = torch.full((B,), KV - 1, dtype=torch.int64, device="cuda")
        self.kv_starts = self.seq_starts * KV
        self.tok = torch.zeros(B, dtype=torch.int64, device="cuda")
        st = capi.StepC()
        st.token_ids, st.seq_starts, st.kv_starts = self.tok.data_ptr(), self.seq_starts.data_ptr(), self.kv_starts.data_ptr()
        st.cache_indices, st.start_pos = self.page_list.data_ptr(), self.start_pos.data_ptr()
        st.num_tokens, st.batch, st.decoding_batches = B, B, B
        st.max_seq_len, st.max_kv_len, st.max_pages = 1, KV, pp
        self.st = st
        self.ws = torch.empty(lib.b2llm_attention_workspace_size(B, HQ, D), dtype=torch.uint8, device="cuda")

    def run(self, stream, layer):
        rc = lib.b2llm_op_attention(C.c_void_p(stream.cuda_stream), _ptr(self.qkv), C.byref(self.st), HQ, C.byref(self.geom),
                                    layer % L, _ptr(self.cache), _ptr(self.scale), _ptr(self.ws), _ptr(self.out), 2)
        assert rc == 0, lib.b2llm_last_error()


class Gemms:
    """o-proj -> (norm+quant) -> gate_up+SwiGLU -> quant -> down -> (norm+quant) -> qkv: the GEMM side of one layer"""
    SETS = 2

    def __init__(self, M):
        self.M = M
        i8 = lambda *s: torch.randint(-127, 128, s, dtype=torch.int8, device="cuda")
        self.w = [dict(o=i8(HID, HID), gu=i8(2 * INTER, HID), dn=i8(HID, INTER), qkv=i8(3 * HID, HID)) for _ in range(self.SETS)]
        self.ws = {k: torch.full((n,), 1e-3, dtype=torch.float32, device="cuda") for k, n in
                   dict(o=HID, gu=2 * INTER, dn=HID, qkv=3 * HID).items()}
        self.a_h, self.a_i = i8(M, HID), i8(M, INTER)
        self.sa = torch.full((M,), 1e-2, dtype=torch.float32, device="cuda")
        self.x = torch.zeros((M, HID), dtype=torch.float16, device="cuda")
        self.act = torch.zeros((M, INTER), dtype=torch.float16, device="cuda")
        self.qkv = torch.zeros((M, 3 * HID), dtype=torch.float16, device="cuda")
        self.gamma = torch.ones((HID,), dtype=torch.float16, device="cuda")

    def run(self, stream, layer):
        s = C.c_void_p(stream.cuda_stream)
        w = self.w[layer % self.SETS]
        M = self.M

        def gemm(a, wk, N, K, epi, out):
            rc = lib.b2llm_op_gemm_w8a8(s, _ptr(a), _ptr(self.sa), _ptr(w[wk]), _ptr(self.ws[wk]), M, N, K, epi, _ptr(out), 0)
            assert rc == 0, lib.b2llm_last_error()

        def normq():
            rc = lib.b2llm_op_rmsnorm_quant(s, _ptr(self.x), None, _ptr(self.gamma), 1e-5, M, HID, _ptr(self.a_h), _ptr(self.sa), None)
            assert rc == 0, lib.b2llm_last_error()

        gemm(self.a_h, "o", HID, HID, 1, self.x)
        normq()
        gemm(self.a_h, "gu", 2 * INTER, HID, 2, self.act)
        rc = lib.b2llm_op_quant_rows(s, _ptr(self.act), M, INTER, _ptr(self.a_i), _ptr(self.sa))
        assert rc == 0, lib.b2llm_last_error()
        gemm(self.a_i, "dn", HID, INTER, 1, self.x)
        normq()
        gemm(self.a_h, "qkv", 3 * HID, HID, 0, self.qkv)


def timed(fn, iters=18):
    """fn(i) enqueues one iteration (on whatever streams); returns ms per iteration, device time over all streams"""
    for i in range(4):
        fn(i)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    base = torch.cuda.current_stream()
    t0.record(base)
    for st in STREAMS:
        st.wait_event(t0)
    for i in range(iters):
        fn(i)
    for st in STREAMS:
        e = torch.cuda.Event()
        e.record(st)
        base.wait_event(e)
    t1.record(base)
    torch.cuda.synchronize()
    return t0.elapsed_time(t1) / iters


lo, hi = torch.cuda.Stream(priority=0), torch.cuda.Stream(priority=-1)
STREAMS = [lo, hi]
H = BFULL // 2
attn_full, attn_half = Attn(BFULL), Attn(H)
g_full, g_half = Gemms(BFULL), Gemms(H)

a_full = timed(lambda i: attn_full.run(lo, i))
g_full_ms = timed(lambda i: g_full.run(lo, i))
serial = timed(lambda i: (attn_full.run(lo, i), g_full.run(lo, i)))
a_half = timed(lambda i: attn_half.run(lo, i))
g_half_ms = timed(lambda i: g_half.run(lo, i))
print(f"full batch {BFULL}: attention {a_full:.4f} ms, GEMM chain {g_full_ms:.4f} ms, serial layer {serial:.4f} ms")
print(f"half batch {H}: attention {a_half:.4f} ms, GEMM chain {g_half_ms:.4f} ms (sum {a_half + g_half_ms:.4f}, max {max(a_half, g_half_ms):.4f})")
for name, sa, sg in (("gemm high priority", lo, hi), ("attention high priority", hi, lo), ("equal priority", lo, torch.cuda.Stream())):
    if sg not in STREAMS:
        STREAMS.append(sg)
    co = timed(lambda i: (attn_half.run(sa, i), g_half.run(sg, i)))
    print(f"co-run [{name}]: {co:.4f} ms per half-layer -> 2 x = {2 * co:.4f} ms vs serial {serial:.4f} ms "
          f"({(1 - 2 * co / serial) * 100:+.1f} % saved)")
