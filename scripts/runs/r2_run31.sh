#!/bin/bash
# NOTE: B2LLM_W4_DIAG was a temporary diagnostic switch of gemm_w4_kernel (skip conversion / MMAs / loads, per-role clock64
# counters); it was removed again after these runs.  Kept as the record of how profiles/r2_gemm_w4_transposed.txt was measured.
# run 31: TEMPORARY in-kernel cycle counters of the W4A16 kernel's roles (B2LLM_W4_DIAG bit 16), gate_up shape, CTA 0
mkdir -p gpurun_out
export FUSED_ONLY=1 SHAPES=gate_up
: > gpurun_out/run31_w4_prof.txt
for d in 16 31 19 17 18; do
  echo "## B2LLM_W4_DIAG=$d" >> gpurun_out/run31_w4_prof.txt
  B2LLM_W4_DIAG=$d timeout 120 python scripts/gemm_w4_bench.py 2>&1 | tail -7 >> gpurun_out/run31_w4_prof.txt
done
echo "## MT=1 forced, diag 16" >> gpurun_out/run31_w4_prof.txt
B2LLM_W4_MT=1 B2LLM_W4_DIAG=16 timeout 120 python scripts/gemm_w4_bench.py 2>&1 | tail -7 >> gpurun_out/run31_w4_prof.txt
cat gpurun_out/run31_w4_prof.txt
