#!/bin/bash
# run 41: per-CTA timeline of one decode-attention launch at the headline shape and at the 70B / TP 8 per-rank shape
mkdir -p gpurun_out
: > gpurun_out/run41_attn_trace.txt
timeout 200 python scripts/attn_trace.py >> gpurun_out/run41_attn_trace.txt 2>&1
B=256 H=8 HKV=1 KV=8192 timeout 200 python scripts/attn_trace.py >> gpurun_out/run41_attn_trace.txt 2>&1
B=1024 H=4 HKV=4 KV=512 timeout 200 python scripts/attn_trace.py >> gpurun_out/run41_attn_trace.txt 2>&1
cat gpurun_out/run41_attn_trace.txt
timeout 300 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "attention" -x 2>&1 | tail -3
