#!/bin/bash
# round 1, run 14 (the last ~8 GPU minutes): BASELINE config 5 -- one whole prefill step through the engine
# (prefill tokens/s, attention / GEMM split) and the tensor-pipe utilisation of the W8A8 GEMMs at M = 8192.
set -x
mkdir -p gpurun_out
SEQS=16 LEN=4096 timeout 150 python scripts/prefill_step_bench.py > gpurun_out/prefill14_16x4096.json 2> gpurun_out/prefill14.err; echo "rc=$?"; cat gpurun_out/prefill14_16x4096.json | cut -c1-900; tail -3 gpurun_out/prefill14.err | cut -c1-300
timeout 120 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum --clock-control none -k regex:"gemm_tc2" -s 4 -c 16 --csv --log-file gpurun_out/gemm_M8192_tensor_pipe14.csv python scripts/gemm_bench.py 8192 > gpurun_out/ncu_gemm14.log 2>&1; echo "ncu rc=$?"; grep -c gemm_tc2 gpurun_out/gemm_M8192_tensor_pipe14.csv
SEQS=64 LEN=4096 timeout 170 python scripts/prefill_step_bench.py > gpurun_out/prefill14_64x4096.json 2>> gpurun_out/prefill14.err; echo "rc=$?"; cat gpurun_out/prefill14_64x4096.json | cut -c1-900; tail -3 gpurun_out/prefill14.err | cut -c1-300
