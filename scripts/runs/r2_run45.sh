#!/bin/bash
# run 45: 70B / TP 8 per-rank attention shape (B 256, 8 q heads on 1 kv head, kv 8192): single-wave plans vs the default 13 x 1
mkdir -p gpurun_out
: > gpurun_out/run45_attn_70b_plans.txt
export B=256 H=8 HKV=1 KV=8192
timeout 100 python scripts/attn_bench.py >> gpurun_out/run45_attn_70b_plans.txt 2>&1
for plan in "6 1" "5 1" "4 1" "3 2" "6 2" "12 1" "14 1"; do
  set -- $plan
  B2LLM_ATTN_SPLITS=$1 B2LLM_ATTN_WARPS=$2 timeout 100 python scripts/attn_bench.py >> gpurun_out/run45_attn_70b_plans.txt 2>&1
done
cut -c1-260 gpurun_out/run45_attn_70b_plans.txt
