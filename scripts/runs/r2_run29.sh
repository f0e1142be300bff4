#!/bin/bash
# NOTE: B2LLM_W4_DIAG was a temporary diagnostic switch of gemm_w4_kernel (skip conversion / MMAs / loads, per-role clock64
# counters); it was removed again after these runs.  Kept as the record of how profiles/r2_gemm_w4_transposed.txt was measured.
# run 29+: W4A16 after per-warp barrier arrival, 19-instruction nibble expansion, vectorised scale loads: correctness, then the diag sweep again
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -m gpu -k "w4" -x 2>&1 | tail -5
export FUSED_ONLY=1
: > gpurun_out/run30_w4_diag.txt
for d in 0 1 2 3 15; do
  echo "## B2LLM_W4_DIAG=$d" >> gpurun_out/run30_w4_diag.txt
  B2LLM_W4_DIAG=$d timeout 120 python scripts/gemm_w4_bench.py >> gpurun_out/run30_w4_diag.txt 2>&1 || echo "rc=$?" >> gpurun_out/run30_w4_diag.txt
done
cat gpurun_out/run30_w4_diag.txt
