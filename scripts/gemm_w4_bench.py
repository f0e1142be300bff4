#!/usr/bin/env python
"""W4A16 GEMM microbenchmark at the per-rank decode shapes of BASELINE config 4 (LLaMA-2-70B, TP=8, M = 256):
fused tcgen05 kernel (nibbles -> converter warps -> UMMA) vs the two-kernel fallback (dequant to fp16 scratch, then the
fp16 tcgen05 GEMM).  Prints us, TFLOP/s and packed-weight GB/s."""
import ctypes as C
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import b200_import  # noqa: E402

b200_import.load()
from ppl_llm_serving_b200 import capi  # noqa: E402
from ppl_llm_serving_b200.engine import _ptr  # noqa: E402

lib = capi.load_library()
M = int(sys.argv[1]) if len(sys.argv) > 1 else 256
shapes = [("qkv", 1280, 8192, 0), ("o", 8192, 1024, 1), ("gate_up", 7168, 8192, 2), ("down", 8192, 3584, 1)]
import os  # noqa: E402
if os.environ.get("SHAPES"):          # e.g. SHAPES=gate_up,down ; FUSED_ONLY=1 skips the two-kernel fallback
    shapes = [s for s in shapes if s[0] in os.environ["SHAPES"].split(",")]
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for name, N, K, epi in shapes:
    copies = max(2, int(300e6 // (N * K // 2)) + 1)
    a = torch.randn((M, K), dtype=torch.float16, device="cuda")
    packs = [torch.randint(0, 256, (N, K // 2), dtype=torch.uint8, device="cuda") for _ in range(copies)]
    scale = (torch.rand((N, K // 128), device="cuda") * 0.01).to(torch.float16)
    out = torch.zeros((M, N if epi != 2 else N // 2), dtype=torch.float16, device="cuda")
    scratch = torch.empty((N, K), dtype=torch.float16, device="cuda")

    def fused(i):
        assert lib.b2llm_op_gemm_w4a16(sp, _ptr(a), _ptr(packs[i % copies]), _ptr(scale), M, N, K, epi, _ptr(out)) == 0, lib.b2llm_last_error()

    def twostep(i):
        assert lib.b2llm_op_dequant_w4(sp, _ptr(packs[i % copies]), _ptr(scale), N, K, _ptr(scratch)) == 0
        assert lib.b2llm_op_gemm_f16(sp, _ptr(a), _ptr(scratch), M, N, K, epi, _ptr(out), 0, 0) == 0, lib.b2llm_last_error()

    for label, fn in ((("fused", fused),) if os.environ.get("FUSED_ONLY") else (("fused", fused), ("dequant+f16", twostep))):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"M={M} {name:8s} N={N:5d} K={K:5d} {label:12s} {us:8.1f} us  {2.0 * M * N * K / us / 1e6:7.1f} TFLOP/s  "
              f"{N * K / 2 / us / 1e3:7.1f} GB/s packed", flush=True)
