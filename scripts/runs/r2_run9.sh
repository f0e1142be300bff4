#!/bin/bash
# round 2, run 9 (1 x B200): whole GPU suite with the tcgen05 prefill kernel as the default, config 5 (prefill step) with
# both prefill kernels, then the ncu evidence (scripts/runs/r2_run7.sh)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2_9_all.log 2>&1; echo "rc=$?" >> gpurun_out/r2_9_all.log; tail -8 gpurun_out/r2_9_all.log | cut -c1-400
IMPL=6 SEQS=8 timeout 120 python scripts/prefill_bench.py > gpurun_out/r2_9_prefill_bench_tc.txt 2>&1; tail -1 gpurun_out/r2_9_prefill_bench_tc.txt
IMPL=8 SEQS=8 timeout 120 python scripts/prefill_bench.py > gpurun_out/r2_9_prefill_bench_mma.txt 2>&1; tail -1 gpurun_out/r2_9_prefill_bench_mma.txt
SEQS=16 timeout 200 python scripts/prefill_step_bench.py > gpurun_out/r2_9_prefill_step_16_tc.json 2> gpurun_out/r2_9_prefill_step.err; cut -c1-700 gpurun_out/r2_9_prefill_step_16_tc.json
B2LLM_PREFILL_IMPL=mma SEQS=16 timeout 200 python scripts/prefill_step_bench.py > gpurun_out/r2_9_prefill_step_16_mma.json 2>> gpurun_out/r2_9_prefill_step.err; cut -c1-700 gpurun_out/r2_9_prefill_step_16_mma.json
SEQS=64 timeout 300 python scripts/prefill_step_bench.py > gpurun_out/r2_9_prefill_step_64_tc.json 2>> gpurun_out/r2_9_prefill_step.err; cut -c1-700 gpurun_out/r2_9_prefill_step_64_tc.json
bash scripts/runs/r2_run7.sh
