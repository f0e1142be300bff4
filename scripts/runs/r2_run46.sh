#!/bin/bash
# run 46: single-wave rule in the decode attention planner: attention + generation tests, then the shapes where it applies
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_golden_gpu.py -q -m gpu -k "attention or generation or golden or kv_budget" -x 2>&1 | tail -3
: > gpurun_out/run46_attn_single_wave.txt
B=256 H=8 HKV=1 KV=8192 timeout 100 python scripts/attn_bench.py >> gpurun_out/run46_attn_single_wave.txt 2>&1
B=256 H=8 HKV=1 KV=8192 B2LLM_ATTN_SPLITS=13 B2LLM_ATTN_WARPS=1 timeout 100 python scripts/attn_bench.py >> gpurun_out/run46_attn_single_wave.txt 2>&1
B=16 H=32 HKV=32 KV=4096 timeout 100 python scripts/attn_bench.py >> gpurun_out/run46_attn_single_wave.txt 2>&1
B=16 H=32 HKV=32 KV=4096 B2LLM_ATTN_SPLITS=5 B2LLM_ATTN_WARPS=2 timeout 100 python scripts/attn_bench.py >> gpurun_out/run46_attn_single_wave.txt 2>&1
B=128 H=8 HKV=1 KV=8192 timeout 100 python scripts/attn_bench.py >> gpurun_out/run46_attn_single_wave.txt 2>&1
B=128 H=8 HKV=1 KV=8192 B2LLM_ATTN_SPLITS=7 B2LLM_ATTN_WARPS=4 timeout 100 python scripts/attn_bench.py >> gpurun_out/run46_attn_single_wave.txt 2>&1
B=256 H=8 HKV=1 KV=8192 timeout 100 python scripts/attn_trace.py >> gpurun_out/run46_attn_single_wave.txt 2>&1
cut -c1-260 gpurun_out/run46_attn_single_wave.txt | grep -v "active CTAs per"
