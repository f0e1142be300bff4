#!/bin/bash
# run 36: transposed W4 kernel with the three-pass direct epilogue: W4 tests, then rule vs 1 vs 2 k-slices
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -m gpu -k "w4" -x 2>&1 | tail -4
export FUSED_ONLY=1
: > gpurun_out/run36_w4t_epi.txt
for sk in 0 1 2; do
  echo "## B2LLM_W4_SPLITK=$sk (0 = rule)" >> gpurun_out/run36_w4t_epi.txt
  B2LLM_W4_SPLITK=$sk timeout 120 python scripts/gemm_w4_bench.py 256 >> gpurun_out/run36_w4t_epi.txt 2>&1 || echo "rc=$?" >> gpurun_out/run36_w4t_epi.txt
done
echo "## 7B / 13B-like shapes at M = 256 (no split-K: direct epilogue)" >> gpurun_out/run36_w4t_epi.txt
cat gpurun_out/run36_w4t_epi.txt
