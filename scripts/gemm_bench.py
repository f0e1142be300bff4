#!/usr/bin/env python
"""W8A8 GEMM microbenchmark at the decode shapes of LLaMA-2-7B (M = running batch): single-CTA tcgen05 kernel
(impl 2) vs CTA-pair kernel (impl 3) vs mma.sync baseline (impl 1).  CUDA events, L2 flushed between runs by
rotating over enough distinct weight copies.  Prints one line per (shape, impl): us, PFLOP/s (int8 ops)."""
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
M = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
shapes = [("qkv", 12288, 4096, 0), ("o", 4096, 4096, 1), ("gate_up", 22016, 4096, 2), ("down", 4096, 11008, 1)]
stream = torch.cuda.current_stream()
sp = C.c_void_p(stream.cuda_stream)
for name, N, K, epi in shapes:
    copies = max(2, int(400e6 // (N * K)) + 1)   # > 3x L2 worth of weights in rotation
    a = torch.randint(-127, 128, (M, K), dtype=torch.int8, device="cuda")
    ws = [torch.randint(-127, 128, (N, K), dtype=torch.int8, device="cuda") for _ in range(copies)]
    sa = torch.rand(M, device="cuda") * 0.01
    sw = torch.rand(N, device="cuda") * 0.001
    out = torch.zeros((M, N if epi != 2 else N // 2), dtype=torch.float16, device="cuda")
    for impl in (1, 2, 3):
        def run(i):
            rc = lib.b2llm_op_gemm_w8a8(sp, _ptr(a), _ptr(sa), _ptr(ws[i % copies]), _ptr(sw), M, N, K, epi, _ptr(out), impl)
            assert rc == 0, lib.b2llm_last_error()
        for i in range(3):
            run(i)
        torch.cuda.synchronize()
        reps = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            run(i)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        print(f"M={M} {name:8s} N={N:6d} K={K:6d} impl={impl} {us:8.1f} us  {2.0 * M * N * K / us / 1e9:6.3f} PFLOP/s", flush=True)
