#!/bin/bash
# round 2, run 8 (2 x B200): join kernel with all peer loads in flight; wave-aware split plan of decode attention;
# the two-threads-per-row tcgen05 prefill kernel
mkdir -p gpurun_out
B2LLM_TEST_EXPERIMENTAL=1 timeout 200 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "prefill_tcgen05_experimental" > gpurun_out/r2_8_prefill_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2_8_prefill_tc.log; tail -8 gpurun_out/r2_8_prefill_tc.log | cut -c1-300
IMPL=6 SEQS=8 timeout 120 python scripts/prefill_bench.py > gpurun_out/r2_8_prefill_bench_tc.txt 2>&1; tail -2 gpurun_out/r2_8_prefill_bench_tc.txt
timeout 400 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "attention" > gpurun_out/r2_8_attn.log 2>&1; echo "rc=$?" >> gpurun_out/r2_8_attn.log; tail -5 gpurun_out/r2_8_attn.log | cut -c1-300
timeout 600 python -m pytest tests/test_tp_gpu.py -x -q -m gpu > gpurun_out/r2_8_tp.log 2>&1; echo "rc=$?" >> gpurun_out/r2_8_tp.log; tail -5 gpurun_out/r2_8_tp.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_8_bench_n2.json 2> gpurun_out/r2_8_bench_n2.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_8_bench_n2.err | cut -c1-300
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_8_bench_n2.json"))
print("replica ms", round(d["ms_per_step"], 2), d["clocks"])
for r in d["config"]["tp"]["runs"]: print("  run", r.get("model"), r.get("ms_per_step"), r.get("device_ms_by_class_per_step"), r.get("fused_join_us_per_call"), r.get("error"))
PY
