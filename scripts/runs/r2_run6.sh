#!/bin/bash
# round 2, run 6 (N x B200): bench.py under torchrun at N = $1 -- replicas + the TP = N leg (parity gate, 7B strong scaling,
# 13B W8A8 at N = 4, 70B GQA W4A16 + the literal 7B 1024 x 2048 at N = 8)
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_6_bench_n$N.json 2> gpurun_out/r2_6_bench_n$N.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_6_bench_n$N.err | cut -c1-600
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2_6_bench_n$N.json"))
    print("replicas", round(d["value"]), "tok/s", round(d["ms_per_step"], 2), "ms frac", round(d["config"]["step_roofline"]["frac_of_hbm_roofline"], 3), d["clocks"])
    tp = d["config"]["tp"]
    for g in tp["parity_gate"]: print("  gate", g)
    for r in tp["runs"]: print("  run", json.dumps(r)[:1500])
    print("  ", tp.get("limiting_collective"))
except Exception as e:
    print("unreadable", e)
PY
