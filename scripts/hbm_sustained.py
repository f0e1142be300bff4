#!/usr/bin/env python
"""Sustained vs burst HBM copy bandwidth on this GPU (context for roofline.frac: MEASURED_PEAKS.json:hbm_gbs is a
best-of-10 burst; the decode step streams ~170 GB per step for tens of ms under the 1 kW power cap).
b.copy_(a) over 1 Gi bf16 elements (2 GiB read + 2 GiB write per copy): best single copy, then the average over a
3-second back-to-back loop."""
import time

import torch

n = 1 << 30
a = torch.empty(n, dtype=torch.bfloat16, device="cuda").normal_()
b = torch.empty_like(a)
bytes_per = 2 * n * 2
for _ in range(3):
    b.copy_(a)
torch.cuda.synchronize()
best = 0.0
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); b.copy_(a); e1.record(); torch.cuda.synchronize()
    best = max(best, bytes_per / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    time.sleep(0.2)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 0
t0 = time.perf_counter()
e0.record()
while time.perf_counter() - t0 < 3.0:
    for _ in range(20):
        b.copy_(a)
    reps += 20
e1.record()
torch.cuda.synchronize()
sus = reps * bytes_per / (e0.elapsed_time(e1) * 1e-3) / 1e9
# read-only stream (the decode attention is ~100 % reads): sum reduction over the same buffer
for _ in range(2):
    a.view(torch.int16).sum()
torch.cuda.synchronize()
e0.record()
for _ in range(40):
    a.view(torch.int16).sum()
e1.record()
torch.cuda.synchronize()
rd = 40 * n * 2 / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(f"hbm copy burst (best of 10, idle between) {best:.0f} GB/s | sustained 3 s loop {sus:.0f} GB/s | read-only torch sum {rd:.0f} GB/s")
