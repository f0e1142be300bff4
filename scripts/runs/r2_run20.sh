#!/bin/bash
# round 2, run 20 (1 x B200): rehearsal of the driver's round-end sequence -- whole GPU suite, smoke(), the reference arm,
# the default bench -- plus the scheduler steady state through the reference's own generator (SURVEY 8d "2c")
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_20_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_20_all.log; tail -6 gpurun_out/r2_20_all.log | cut -c1-400
timeout 200 python __graft_entry__.py smoke > gpurun_out/r2_20_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_20_smoke.log | cut -c1-300
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/r2_20_bench_reference.json 2> gpurun_out/r2_20_bench_reference.err; echo "reference rc=$?"; cut -c1-700 gpurun_out/r2_20_bench_reference.json
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_20_bench.json 2> gpurun_out/r2_20_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_20_bench.json"))
print(round(d["value"]), "tok/s", round(d["ms_per_step"], 3), "ms  frac", round(d["config"]["step_roofline"]["frac_of_hbm_roofline"], 4), "attn", round(d["roofline"]["avg_launch_ms"], 4), round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"], {k: round(v, 3) for k, v in d["config"]["device_ms_by_class_per_step"].items() if k != "note"}, d["config"].get("alt_shapes"), d.get("cpu_baseline", {}).get("value"), d["gpu_launches"])
PY
timeout 900 python scripts/steady_state.py > gpurun_out/r2_20_steady_state.txt 2>&1; echo "steady rc=$?"; tail -3 gpurun_out/r2_20_steady_state.txt | cut -c1-600
