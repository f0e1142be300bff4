#!/bin/bash
# run 37: fixed vs per-k-block cost of the one-wave W8A8 GEMMs (o / down shapes), fp16 vs residual epilogue, PDL on / off
mkdir -p gpurun_out
: > gpurun_out/run37_gemm_overhead.txt
for pdl in 1 0; do
  echo "## B2LLM_PDL=$pdl" >> gpurun_out/run37_gemm_overhead.txt
  B2LLM_PDL=$pdl timeout 200 python scripts/gemm_overhead_probe.py 4096 >> gpurun_out/run37_gemm_overhead.txt 2>&1
done
cat gpurun_out/run37_gemm_overhead.txt
