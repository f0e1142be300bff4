"""Summarise an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list of bench.py per kernel:
launches, total ms, share of the step kernels, average us.  usage: summarize_launches.py launches.csv [steps_in_process]"""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] == "gpu__time_duration.sum":
        rows.append((r["Kernel Name"], float(r["Metric Value"].replace(",", ""))))
LOAD_ONLY = ("synth_fp16", "quant_weight", "interleave_rows", "fill_kv", "Memset", "elementwise", "vectorized")
agg = defaultdict(lambda: [0, 0.0])
for name, ns in rows:
    short = re.sub(r"^(void )?b2llm::(<unnamed>::)?", "", name)
    short = re.sub(r"\(.*$", "", short)
    if any(k in short for k in LOAD_ONLY) or short.startswith("at::") or "at::native" in name:
        continue
    agg[short][0] += 1
    agg[short][1] += ns / 1e6
total = sum(v[1] for v in agg.values())
print(f"{'kernel':52s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:52s} {n:8d} {ms:10.2f} {100 * ms / total:6.1f}% {1e3 * ms / n:9.1f}")
print(f"{'(step kernels total)':52s} {sum(v[0] for v in agg.values()):8d} {total:10.2f}")
if steps:
    print(f"# per step (1/{steps} of the above): {total / steps:.2f} ms")
