#!/usr/bin/env python
"""W8A8 GEMM, one wave (N = 4096 -> 64 pair tiles on 74 CTA pairs), M = 1024: fixed cost vs per-k-block cost.
Times the CTA-pair kernel over K = 1024 .. 16384 with the fp16 (0) and the residual (1) epilogue; a straight-line fit over K
gives (fixed us, us per 128-byte k-block).  Is the residual epilogue's read-modify-write exposed?  (round 2 run 37)"""
import ctypes as C
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
M, N = 1024, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sp = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for epi in (0, 1):
    pts = []
    for K in (1024, 2048, 4096, 8192, 16384):
        copies = max(2, int(400e6 // (N * K)) + 1)
        a = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
        ws = [torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda") for _ in range(copies)]
        sa = torch.rand(M, device="cuda") * 0.01
        sw = torch.rand(N, device="cuda") * 0.001
        out = torch.zeros((M, N), dtype=torch.float16, device="cuda")

        def run(i):
            assert lib.b2llm_op_gemm_w8a8(sp, _ptr(a), _ptr(sa), _ptr(ws[i % copies]), _ptr(sw), M, N, K, epi, _ptr(out), 3) == 0

        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        reps = 30
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        pts.append((K, us))
        del ws
    ks, us = np.array([p[0] for p in pts], float), np.array([p[1] for p in pts])
    slope, icpt = np.polyfit(ks / 128.0, us, 1)
    print(f"N={N} epilogue {epi}: " + "  ".join(f"K={k}: {u:.1f} us" for k, u in pts) +
          f"  -> fixed {icpt:.1f} us + {slope * 1e3:.0f} ns per k-block (tensor floor at 1.9 GHz: 4 x 128 cycles = 270 ns)", flush=True)
