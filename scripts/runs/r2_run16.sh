#!/bin/bash
# round 2, run 16 (1 x B200): A/B on one box -- register-resident row kernels at decode sizes; default for reference
mkdir -p gpurun_out
for v in default reg default reg; do
  if [ $v = reg ]; then export B2LLM_ROW_KERNELS=reg; else unset B2LLM_ROW_KERNELS; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-alt > gpurun_out/r2_16_bench_$v.json 2> gpurun_out/r2_16_bench_$v.err
  python - <<PY
import json
d = json.load(open("gpurun_out/r2_16_bench_$v.json"))
print("$v", round(d["value"]), "tok/s", round(d["ms_per_step"], 3), "ms", "attn", round(d["roofline"]["avg_launch_ms"], 4), d["clocks"]["sm_mhz"], {k: round(v, 3) for k, v in d["config"]["device_ms_by_class_per_step"].items() if k != "note"})
PY
done
