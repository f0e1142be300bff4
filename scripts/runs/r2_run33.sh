#!/bin/bash
# NOTE: B2LLM_W4_PAIR selected an in-kernel two-slice reduction that was measured here, found slower and removed again
# (profiles/r2_gemm_w4_transposed.txt section 5).
# run 33: transposed W4A16 kernel with the in-kernel two-slice reduction (clusters of two, DSMEM): correctness, then timing
# with and without it (B2LLM_W4_PAIR=0 -> fp32 scratch + reduce kernel)
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -m gpu -k "w4" -x 2>&1 | tail -15
echo "pytest rc=${PIPESTATUS[0]}"
export FUSED_ONLY=1
: > gpurun_out/run33_w4t_pair.txt
for pair in 1 0; do
  for M in 256 128; do
    echo "## B2LLM_W4_PAIR=$pair M=$M" >> gpurun_out/run33_w4t_pair.txt
    B2LLM_W4_PAIR=$pair timeout 120 python scripts/gemm_w4_bench.py $M >> gpurun_out/run33_w4t_pair.txt 2>&1 || echo "rc=$?" >> gpurun_out/run33_w4t_pair.txt
  done
done
cat gpurun_out/run33_w4t_pair.txt
