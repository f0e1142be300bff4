#!/bin/bash
# run 42: split-KV merge kernel with parallel loads: attention tests (merge is bit-identical by construction), then the
# split shapes in steady state (attn_bench) and one trace
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -q -m gpu -k "attention or generation" -x 2>&1 | tail -3
: > gpurun_out/run42_attn_merge.txt
B=256 H=8 HKV=1 KV=8192 timeout 200 python scripts/attn_bench.py >> gpurun_out/run42_attn_merge.txt 2>&1
B=1024 H=4 HKV=4 KV=512 timeout 200 python scripts/attn_bench.py >> gpurun_out/run42_attn_merge.txt 2>&1
B=1024 H=4 HKV=4 KV=2048 timeout 200 python scripts/attn_bench.py >> gpurun_out/run42_attn_merge.txt 2>&1
B=256 H=8 HKV=1 KV=8192 timeout 200 python scripts/attn_trace.py >> gpurun_out/run42_attn_merge.txt 2>&1
cat gpurun_out/run42_attn_merge.txt
