#!/bin/bash
# NOTE: B2LLM_W4_DIAG was a temporary diagnostic switch of gemm_w4_kernel (skip conversion / MMAs / loads, per-role clock64
# counters); it was removed again after these runs.  Kept as the record of how profiles/r2_gemm_w4_transposed.txt was measured.
# run 28: where does the W4A16 kernel's time go?  TEMPORARY diag knob (B2LLM_W4_DIAG) + one ncu --set full capture of gate_up
mkdir -p gpurun_out
export FUSED_ONLY=1
: > gpurun_out/run28_w4_diag.txt
for d in 0 1 2 3 4 8 12 13 15; do
  echo "## B2LLM_W4_DIAG=$d" >> gpurun_out/run28_w4_diag.txt
  B2LLM_W4_DIAG=$d timeout 120 python scripts/gemm_w4_bench.py >> gpurun_out/run28_w4_diag.txt 2>&1 || echo "rc=$?" >> gpurun_out/run28_w4_diag.txt
done
cat gpurun_out/run28_w4_diag.txt
SHAPES=gate_up timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_w4_kernel -s 6 -c 1 \
  -o gpurun_out/run28_w4_gateup -f python scripts/gemm_w4_bench.py > gpurun_out/run28_ncu.log 2>&1
tail -3 gpurun_out/run28_ncu.log; ls -la gpurun_out/*.ncu-rep
