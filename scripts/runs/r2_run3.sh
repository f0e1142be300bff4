#!/bin/bash
# round 2, run 3 (1 x B200): fp16 KV cache mode (cache_quant_bit 0) -- op tests, engine tests; then the whole GPU suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "fp16_cache or rope_kv" > gpurun_out/r2_3_ops_fp16kv.log 2>&1; echo "rc=$?" >> gpurun_out/r2_3_ops_fp16kv.log; tail -25 gpurun_out/r2_3_ops_fp16kv.log | cut -c1-400
timeout 300 python -m pytest tests/test_engine_gpu.py -x -q -m gpu -k "fp16_cache" > gpurun_out/r2_3_engine_fp16kv.log 2>&1; echo "rc=$?" >> gpurun_out/r2_3_engine_fp16kv.log; tail -25 gpurun_out/r2_3_engine_fp16kv.log | cut -c1-400
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/r2_3_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_3_all.log; tail -30 gpurun_out/r2_3_all.log | cut -c1-400
