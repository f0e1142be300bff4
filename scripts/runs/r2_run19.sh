#!/bin/bash
# round 2, run 19 (2 x B200): join with acq_rel fences instead of sequentially consistent ones
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -x -q -m gpu -k "fused or failmap" > gpurun_out/r2_19_tp.log 2>&1; echo "rc=$?" >> gpurun_out/r2_19_tp.log; tail -4 gpurun_out/r2_19_tp.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_19_bench_n2.json 2> gpurun_out/r2_19_bench_n2.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_19_bench_n2.json"))
print("replica ms", round(d["ms_per_step"], 2), d["clocks"])
for r in d["config"]["tp"]["runs"]: print("  run", r.get("model"), r.get("ms_per_step"), r.get("device_ms_by_class_per_step"), r.get("fused_join_us_per_call"), r.get("error"))
PY
