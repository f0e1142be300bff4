#!/bin/bash
# round 2, run 2 (2 x B200): tensor parallelism -- vocab-parallel head, TP parity tests, bench.py TP leg at N = 2;
# plus the new GEMM shapes (K tail), the penalty continuity rule and the default LOADER 3 on the op tests
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_tp_gpu.py -x -q -m gpu > gpurun_out/r2_2_tp.log 2>&1; echo "rc=$?" >> gpurun_out/r2_2_tp.log; tail -15 gpurun_out/r2_2_tp.log | cut -c1-300
timeout 400 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k "tensor_parallel_2" > gpurun_out/r2_2_tp_host.log 2>&1; echo "rc=$?" >> gpurun_out/r2_2_tp_host.log; tail -15 gpurun_out/r2_2_tp_host.log | cut -c1-300
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/r2_2_ops.log 2>&1; echo "rc=$?" >> gpurun_out/r2_2_ops.log; tail -15 gpurun_out/r2_2_ops.log | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > gpurun_out/r2_2_bench_n2.json 2> gpurun_out/r2_2_bench_n2.err; echo "bench rc=$?"; tail -5 gpurun_out/r2_2_bench_n2.err | cut -c1-400; cut -c1-3000 gpurun_out/r2_2_bench_n2.json
