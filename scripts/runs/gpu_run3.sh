#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -k "gemm" > gpurun_out/pytest_gemm.log 2>&1; echo "gemm rc=$?" | tee -a gpurun_out/pytest_gemm.log
tail -15 gpurun_out/pytest_gemm.log
timeout 200 python scripts/gemm_bench.py 1024 > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench rc=$?"
cat gpurun_out/bench3.json
B2LLM_GEMM_2CTA=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench3_1cta.json 2>> gpurun_out/bench3.err
cat gpurun_out/bench3_1cta.json
