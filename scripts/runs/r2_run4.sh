#!/bin/bash
# round 2, run 4 (2 x B200): the fused TP residual join (all-reduce + residual + RMSNorm + quant over NVLink peer memory)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tp_gpu.py -x -q -m gpu > gpurun_out/r2_4_tp.log 2>&1; echo "rc=$?" >> gpurun_out/r2_4_tp.log; tail -25 gpurun_out/r2_4_tp.log | cut -c1-400
timeout 400 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k "tensor_parallel_2" > gpurun_out/r2_4_tp_host.log 2>&1; echo "rc=$?" >> gpurun_out/r2_4_tp_host.log; tail -15 gpurun_out/r2_4_tp_host.log | cut -c1-400
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_4_bench_n2.json 2> gpurun_out/r2_4_bench_n2.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_4_bench_n2.err | cut -c1-400
B2LLM_TP_JOIN=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_4_bench_n2_nccl.json 2> gpurun_out/r2_4_bench_n2_nccl.err; echo "bench nccl rc=$?"
python - <<'PY'
import json
for f in ("gpurun_out/r2_4_bench_n2.json", "gpurun_out/r2_4_bench_n2_nccl.json"):
    try:
        d = json.load(open(f))
        print(f, "replica ms", round(d["ms_per_step"], 2), "frac", round(d["config"]["step_roofline"]["frac_of_hbm_roofline"], 3), d["clocks"])
        for g in d["config"]["tp"]["parity_gate"]: print("  gate", g.get("passed"), g.get("logits_rel_err_max"), g.get("error"))
        for r in d["config"]["tp"]["runs"]: print("  run", r.get("model"), r.get("ms_per_step"), r.get("device_ms_by_class_per_step"), r.get("error"))
    except Exception as e:
        print(f, "unreadable", e)
PY
