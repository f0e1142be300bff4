#!/bin/bash
# round 2, run 5 (1 x B200): programmatic dependent launch across the step kernels (+ weight pre-loads before the wait in
# the GEMMs): whole GPU suite with PDL on, then the bench with and without it, back to back on the same box
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_5_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_5_all.log; tail -8 gpurun_out/r2_5_all.log | cut -c1-400
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_5_bench_pdl.json 2> gpurun_out/r2_5_bench_pdl.err; echo "rc=$?"
B2LLM_PDL=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_5_bench_nopdl.json 2> gpurun_out/r2_5_bench_nopdl.err; echo "rc=$?"
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2_5_bench_pdl2.json 2> gpurun_out/r2_5_bench_pdl2.err; echo "rc=$?"
python - <<'PY'
import json
for f in ("pdl", "nopdl", "pdl2"):
    try:
        d = json.load(open(f"gpurun_out/r2_5_bench_{f}.json"))
        print(f, round(d["value"]), "tok/s", round(d["ms_per_step"], 3), "ms  frac", round(d["config"]["step_roofline"]["frac_of_hbm_roofline"], 4),
              "attn", round(d["roofline"]["avg_launch_ms"], 4), round(d["roofline"]["frac"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"],
              {k: round(v, 3) for k, v in d["config"]["device_ms_by_class_per_step"].items() if k != "note"}, d["config"].get("alt_shapes"))
    except Exception as e:
        print(f, "unreadable", e)
PY
