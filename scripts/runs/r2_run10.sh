#!/bin/bash
# round 2, run 10 (1 x B200): whole GPU suite (tcgen05 prefill with hi + lo P as the default; W4A16 GEMM with 8 converter
# warps + split-K), prefill and W4A16 micro-benchmarks, config 5
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_10_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_10_all.log; tail -8 gpurun_out/r2_10_all.log | cut -c1-400
IMPL=6 SEQS=8 timeout 120 python scripts/prefill_bench.py > gpurun_out/r2_10_prefill_bench_tc.txt 2>&1; tail -1 gpurun_out/r2_10_prefill_bench_tc.txt
timeout 200 python scripts/gemm_w4_bench.py > gpurun_out/r2_10_gemm_w4.txt 2>&1; tail -12 gpurun_out/r2_10_gemm_w4.txt
SEQS=16 timeout 200 python scripts/prefill_step_bench.py > gpurun_out/r2_10_prefill_step_16_tc.json 2> gpurun_out/r2_10_prefill_step.err; cut -c1-600 gpurun_out/r2_10_prefill_step_16_tc.json
SEQS=64 timeout 300 python scripts/prefill_step_bench.py > gpurun_out/r2_10_prefill_step_64_tc.json 2>> gpurun_out/r2_10_prefill_step.err; cut -c1-600 gpurun_out/r2_10_prefill_step_64_tc.json
