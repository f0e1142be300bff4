#!/usr/bin/env python
"""Can the tensor-bound W8A8 GEMMs hide under the HBM-bound decode attention?  Launches the attention op (B=1024,
kv_len 512, one layer's worth = 5.4 GB) on one stream and a chain of gate_up GEMMs (M=512) on another, alone and
together, and prints the three times.  B2LLM_GEMM_SMS limits the SMs the persistent GEMM grid occupies."""
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
B, KV, H, D, L, PAGE = 1024, 512, 32, 128, 8, 16
T = B * KV
geom = capi.KvGeomC()
geom.num_layers, geom.num_kv_heads, geom.head_dim, geom.quant_group = L, H, D, 8
geom.cache_layout, geom.cache_mode, geom.page_size, geom.max_tokens = 3, 1, PAGE, T
cache = torch.randint(-127, 128, (L * 2 * H * T * D,), dtype=torch.int8, device="cuda")
scale = torch.full((L * 2 * H * T * D // 8,), 0.01, dtype=torch.float16, device="cuda")
qkv = torch.randn((B, 3 * H * D), dtype=torch.float16, device="cuda")
out = torch.empty((B, H * D), dtype=torch.float16, device="cuda")
rng = np.random.default_rng(0)
pages_per = KV // PAGE
page_list = torch.from_numpy((rng.permutation(B * pages_per).reshape(B, pages_per) * PAGE).astype(np.int64)).cuda()
seq_starts = torch.arange(B + 1, dtype=torch.int64, device="cuda")
start_pos = torch.full((B,), KV - 1, dtype=torch.int64, device="cuda")
kv_starts = seq_starts * KV
tok = torch.zeros(B, dtype=torch.int64, device="cuda")
st = capi.StepC()
st.token_ids, st.seq_starts, st.kv_starts = tok.data_ptr(), seq_starts.data_ptr(), kv_starts.data_ptr()
st.cache_indices, st.start_pos = page_list.data_ptr(), start_pos.data_ptr()
st.num_tokens, st.batch, st.decoding_batches = B, B, B
st.max_seq_len, st.max_kv_len, st.max_pages = 1, KV, pages_per
ws = torch.empty(lib.b2llm_attention_workspace_size(B, H, D), dtype=torch.uint8, device="cuda")

M, N, K = int(os.environ.get("GEMM_M", 512)), 22016, 4096
a = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
wts = [torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda") for _ in range(4)]
sa = torch.rand(M, device="cuda") * 0.01
sw = torch.rand(N, device="cuda") * 0.001
gout = torch.zeros((M, N // 2), dtype=torch.float16, device="cuda")
NG = int(os.environ.get("GEMMS", 12))

s_att, s_gemm = torch.cuda.Stream(), torch.cuda.Stream(priority=-1)


def attn(layer):
    rc = lib.b2llm_op_attention(C.c_void_p(s_att.cuda_stream), _ptr(qkv), C.byref(st), H, C.byref(geom), layer, _ptr(cache),
                                _ptr(scale), _ptr(ws), _ptr(out), 2)
    assert rc == 0


def gemms():
    for i in range(NG):
        rc = lib.b2llm_op_gemm_w8a8(C.c_void_p(s_gemm.cuda_stream), _ptr(a), _ptr(sa), _ptr(wts[i % 4]), _ptr(sw), M, N, K, 2,
                                    _ptr(gout), 0)
        assert rc == 0


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    s_att.wait_event(e0); s_gemm.wait_event(e0)
    for _ in range(reps):
        fn()
    ea, eg = torch.cuda.Event(), torch.cuda.Event()
    ea.record(s_att); eg.record(s_gemm)
    torch.cuda.current_stream().wait_event(ea); torch.cuda.current_stream().wait_event(eg)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


layer = [0]


def only_attn():
    for _ in range(2):
        attn(layer[0] % L); layer[0] += 1


def both_gemm_first():
    gemms(); only_attn()


def both_attn_first():
    only_attn(); gemms()


ta, tg = timed(only_attn), timed(gemms)
tb1, tb2 = timed(both_gemm_first), timed(both_attn_first)
print(f"GEMM_SMS={os.environ.get('B2LLM_GEMM_SMS', 'all')} M={M}: 2x attention {ta:.3f} ms | {NG} gate_up GEMMs {tg:.3f} ms | "
      f"together (gemm launched first) {tb1:.3f} ms, (attention first) {tb2:.3f} ms | sum {ta + tg:.3f}", flush=True)
