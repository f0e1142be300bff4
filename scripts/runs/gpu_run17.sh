#!/bin/bash
# round 1, run 17 (last GPU seconds): the "slim" decode-attention loader (B2LLM_ATTN_SLIM=1: incremental page walk instead
# of an integer division per unit, bare ex2.approx) -- parity tests of the attention op, then the step with and without it.
mkdir -p gpurun_out
B2LLM_ATTN_SLIM=1 timeout 60 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "attention" > gpurun_out/pytest17_slim.log 2>&1; echo "rc=$?" >> gpurun_out/pytest17_slim.log; tail -3 gpurun_out/pytest17_slim.log | cut -c1-200
B2LLM_ATTN_SLIM=1 timeout 70 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench17_slim.json 2> gpurun_out/bench17.err; python -c "
import json; d=json.load(open('gpurun_out/bench17_slim.json')); print('slim', d['value'], d['roofline']['avg_launch_ms'], d['clocks'])"
timeout 70 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench17_base.json 2>> gpurun_out/bench17.err; python -c "
import json; d=json.load(open('gpurun_out/bench17_base.json')); print('base', d['value'], d['roofline']['avg_launch_ms'], d['clocks'])"
B2LLM_ATTN_SLIM=1 timeout 90 python -m pytest tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/pytest17_engine_slim.log 2>&1; echo "rc=$?" >> gpurun_out/pytest17_engine_slim.log; tail -3 gpurun_out/pytest17_engine_slim.log | cut -c1-200
