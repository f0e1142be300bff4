#!/bin/bash
# round 2, run 15 (1 x B200): GEMM epilogue on 8 warps (two per TMEM lane quarter) + vectorised scale loads
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "gemm" > gpurun_out/r2_15_gemm.log 2>&1; echo "rc=$?" >> gpurun_out/r2_15_gemm.log; tail -5 gpurun_out/r2_15_gemm.log | cut -c1-300
timeout 200 python scripts/gemm_bench.py > gpurun_out/r2_15_gemm_w8.txt 2>&1; grep "impl=3\|impl=2" gpurun_out/r2_15_gemm_w8.txt
M=8192 timeout 200 python scripts/gemm_bench.py > gpurun_out/r2_15_gemm_w8_M8192.txt 2>&1; grep "impl=3" gpurun_out/r2_15_gemm_w8_M8192.txt
timeout 600 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/r2_15_engine.log 2>&1; echo "rc=$?" >> gpurun_out/r2_15_engine.log; tail -4 gpurun_out/r2_15_engine.log | cut -c1-300
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_15_bench.json 2> gpurun_out/r2_15_bench.err; echo "rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_15_bench.json"))
print(round(d["value"]), "tok/s", round(d["ms_per_step"], 3), "ms  frac", round(d["config"]["step_roofline"]["frac_of_hbm_roofline"], 4), "attn", round(d["roofline"]["avg_launch_ms"], 4), round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"], {k: round(v, 3) for k, v in d["config"]["device_ms_by_class_per_step"].items() if k != "note"}, d["config"].get("alt_shapes"))
PY
