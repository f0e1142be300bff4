#!/bin/bash
# run 32: transposed W4A16 kernel (weight tile through TMEM): correctness, then A/B against the both-in-smem kernel
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -m gpu -k "w4" -x 2>&1 | tail -15
echo "pytest rc=${PIPESTATUS[0]}"
export FUSED_ONLY=1
: > gpurun_out/run32_w4t.txt
for impl in ts ss; do
  for M in 256 128 64; do
    echo "## B2LLM_W4_IMPL=$impl M=$M" >> gpurun_out/run32_w4t.txt
    B2LLM_W4_IMPL=$impl timeout 120 python scripts/gemm_w4_bench.py $M >> gpurun_out/run32_w4t.txt 2>&1 || echo "rc=$?" >> gpurun_out/run32_w4t.txt
  done
done
cat gpurun_out/run32_w4t.txt
