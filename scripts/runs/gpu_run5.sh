#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_host_cpp.py tests/test_tp_gpu.py -x -q -m gpu -k "attention or generation or config1 or w8a8_7b or reference or tensor_parallel" > gpurun_out/pytest5.log 2>&1; echo "rc=$?" | tee -a gpurun_out/pytest5.log
tail -40 gpurun_out/pytest5.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
cat gpurun_out/bench_n2.json
