#!/bin/bash
# 2-GPU session: TP tests (python one-process-per-GPU and the reference's single-process C++ path), replicas bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_tp_gpu.py tests/test_host_cpp.py -x -q -m gpu -k "tensor_parallel" > gpurun_out/pytest_tp.log 2>&1; echo "tp rc=$?" | tee -a gpurun_out/pytest_tp.log
tail -30 gpurun_out/pytest_tp.log
timeout 200 python scripts/attn_bench.py > gpurun_out/attn_bench.log 2>&1
B2LLM_ATTN_WARPS=2 timeout 200 python scripts/attn_bench.py >> gpurun_out/attn_bench.log 2>&1
B2LLM_ATTN_WARPS=4 timeout 200 python scripts/attn_bench.py >> gpurun_out/attn_bench.log 2>&1
cat gpurun_out/attn_bench.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench2 rc=$?"
cat gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
