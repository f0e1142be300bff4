#!/bin/bash
# run 34: W4 tests on the cleaned-up library, then the split-K rule of the transposed kernel (B2LLM_W4_SPLITK override)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -m gpu -k "w4" -x 2>&1 | tail -4
export FUSED_ONLY=1
: > gpurun_out/run34_w4t_split.txt
for sk in 0 1 2 3 4 5 7; do
  echo "## B2LLM_W4_SPLITK=$sk (0 = rule)" >> gpurun_out/run34_w4t_split.txt
  B2LLM_W4_SPLITK=$sk timeout 120 python scripts/gemm_w4_bench.py 256 >> gpurun_out/run34_w4t_split.txt 2>&1 || echo "rc=$?" >> gpurun_out/run34_w4t_split.txt
done
cat gpurun_out/run34_w4t_split.txt
