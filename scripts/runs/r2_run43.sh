#!/bin/bash
# run 43: decode attention with 4 ring stages per warp (3 units in flight, 10 one-warp CTAs per SM) against the default
# (3 stages, 12 CTAs per SM), same box: steady state (attn_bench) and one per-CTA trace per shape
mkdir -p gpurun_out
: > gpurun_out/run43_attn_nstage.txt
for lib in "" ppl.llm.serving_b200/lib/libb2llm_nstage4.so; do
  export B2LLM_LIB=$lib
  echo "## B2LLM_LIB=${lib:-default (NSTAGE 3)}" >> gpurun_out/run43_attn_nstage.txt
  timeout 200 python scripts/attn_bench.py >> gpurun_out/run43_attn_nstage.txt 2>&1
  B=256 H=8 HKV=1 KV=8192 timeout 200 python scripts/attn_bench.py >> gpurun_out/run43_attn_nstage.txt 2>&1
  B=1024 H=4 HKV=4 KV=512 timeout 200 python scripts/attn_bench.py >> gpurun_out/run43_attn_nstage.txt 2>&1
  B=512 H=10 HKV=10 KV=2640 timeout 200 python scripts/attn_bench.py >> gpurun_out/run43_attn_nstage.txt 2>&1
  timeout 200 python scripts/attn_trace.py >> gpurun_out/run43_attn_nstage.txt 2>&1
  B=256 H=8 HKV=1 KV=8192 timeout 200 python scripts/attn_trace.py >> gpurun_out/run43_attn_nstage.txt 2>&1
done
cut -c1-400 gpurun_out/run43_attn_nstage.txt | grep -v "active CTAs per"
