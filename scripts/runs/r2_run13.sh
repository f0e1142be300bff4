#!/bin/bash
# round 2, run 13 (2 x B200): where do the 25 us of the fused join go?  finer phase clocks (CTA 0's first row)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_13_bench_n2.json 2> gpurun_out/r2_13_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_13_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_13_bench_n2.json"))
print("replica ms", round(d["ms_per_step"], 2), d["clocks"])
for r in d["config"]["tp"]["runs"]: print("  run", r.get("model"), r.get("ms_per_step"), r.get("device_ms_by_class_per_step"), r.get("fused_join_us_per_call"), r.get("error"))
PY
