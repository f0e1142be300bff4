#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -x -q -m gpu > gpurun_out/pytest12.log 2>&1; echo "rc=$?" | tee -a gpurun_out/pytest12.log; tail -6 gpurun_out/pytest12.log | cut -c1-250
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench12.json 2> gpurun_out/bench12.err; cat gpurun_out/bench12.json | cut -c1-400
B2LLM_ROW_KERNELS=legacy timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench12_legacy_rows.json 2>> gpurun_out/bench12.err; cat gpurun_out/bench12_legacy_rows.json | cut -c1-400
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "rmsnorm or quant_rows or rope_kv or attention_decode_mha or (gemm_w8a8_f16_bit_exact and 300) or w4a16_fused and 100" > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?" | tee -a gpurun_out/memcheck.log; tail -8 gpurun_out/memcheck.log | cut -c1-250
