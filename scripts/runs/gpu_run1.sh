#!/bin/bash
# GPU session: parity tests, bench line, ncu launch list, ncu full capture of the top kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt
nproc >> gpurun_out/gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 100 -c 2 -o gpurun_out/prof_attn -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_attn.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 400 -c 4 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_gemm.log 2>&1
ls -la gpurun_out
