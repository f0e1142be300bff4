#!/bin/bash
# round 2, run 17 (1 x B200): row kernels (register-resident from 256 rows) parity; W4A16: two A tiles per weight tile + split-K?
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "rmsnorm or quant_rows" > gpurun_out/r2_17_rows.log 2>&1; echo "rc=$?" >> gpurun_out/r2_17_rows.log; tail -4 gpurun_out/r2_17_rows.log | cut -c1-300
timeout 200 python scripts/gemm_w4_bench.py > gpurun_out/r2_17_gemm_w4_default.txt 2>&1; grep fused gpurun_out/r2_17_gemm_w4_default.txt
B2LLM_W4_MT=2 timeout 200 python scripts/gemm_w4_bench.py > gpurun_out/r2_17_gemm_w4_mt2.txt 2>&1; grep fused gpurun_out/r2_17_gemm_w4_mt2.txt
B2LLM_W4_MT=1 timeout 200 python scripts/gemm_w4_bench.py > gpurun_out/r2_17_gemm_w4_mt1.txt 2>&1; grep fused gpurun_out/r2_17_gemm_w4_mt1.txt
