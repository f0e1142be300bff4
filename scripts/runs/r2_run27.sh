#!/bin/bash
# run 27: golden replay with the final bounds; whole-layer two-stream overlap probe at three GEMM SM budgets
mkdir -p gpurun_out
python -m pytest tests/test_golden_gpu.py -q -m gpu -s > gpurun_out/run27_golden.log 2>&1; echo "golden rc=$?"
grep -E "GOLDEN|passed|failed" gpurun_out/run27_golden.log
: > gpurun_out/run27_overlap.txt
for sms in all 96 64; do
  if [ "$sms" = all ]; then unset B2LLM_GEMM_SMS; else export B2LLM_GEMM_SMS=$sms; fi
  echo "## B2LLM_GEMM_SMS=$sms" >> gpurun_out/run27_overlap.txt
  timeout 240 python scripts/overlap_probe_layer.py >> gpurun_out/run27_overlap.txt 2>&1 || echo "probe rc=$?" >> gpurun_out/run27_overlap.txt
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv >> gpurun_out/run27_overlap.txt
cat gpurun_out/run27_overlap.txt
