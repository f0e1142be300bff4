#!/bin/bash
# round 2, run 11 (1 x B200): W8A8 pair GEMM with L2 prefetch of the weight tiles 16 k-blocks ahead; W4A16 split-K with a
# separate reduce kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/r2_11_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_11_gemm.log; tail -5 gpurun_out/r2_11_gemm.log | cut -c1-300
timeout 200 python scripts/gemm_w4_bench.py > gpurun_out/r2_11_gemm_w4.txt 2>&1; grep fused gpurun_out/r2_11_gemm_w4.txt
timeout 200 python scripts/gemm_bench.py > gpurun_out/r2_11_gemm_w8.txt 2>&1; tail -12 gpurun_out/r2_11_gemm_w8.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-alt > gpurun_out/r2_11_bench.json 2> gpurun_out/r2_11_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_11_bench.json"))
print(round(d["value"]), "tok/s", round(d["ms_per_step"], 3), "ms  frac", round(d["config"]["step_roofline"]["frac_of_hbm_roofline"], 4), "attn", round(d["roofline"]["avg_launch_ms"], 4), round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"], {k: round(v, 3) for k, v in d["config"]["device_ms_by_class_per_step"].items() if k != "note"})
PY
