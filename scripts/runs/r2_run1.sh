#!/bin/bash
# round 2, run 1: bring-up of the two never-run kernels (tcgen05 prefill attention, decode LOADER 3) + config 2b shape
mkdir -p gpurun_out
B2LLM_TEST_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "merged_loader_experimental" > gpurun_out/r2_1_loader3.log 2>&1; echo "rc=$?" >> gpurun_out/r2_1_loader3.log; tail -5 gpurun_out/r2_1_loader3.log | cut -c1-300
B2LLM_TEST_EXPERIMENTAL=1 timeout 150 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "prefill_tcgen05_experimental" > gpurun_out/r2_1_prefill_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2_1_prefill_tc.log; tail -30 gpurun_out/r2_1_prefill_tc.log | cut -c1-300
B2LLM_ATTN_SLIM=2 timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r2_1_bench_loader3.json 2> gpurun_out/r2_1_bench_loader3.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2_1_bench_loader3.json
timeout 120 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/r2_1_bench_base.json 2> gpurun_out/r2_1_bench_base.err; echo "rc=$?"; cut -c1-400 gpurun_out/r2_1_bench_base.json
IMPL=6 SEQS=8 timeout 120 python scripts/prefill_bench.py > gpurun_out/r2_1_prefill_bench_tc.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_1_prefill_bench_tc.txt
IMPL=2 SEQS=8 timeout 120 python scripts/prefill_bench.py > gpurun_out/r2_1_prefill_bench_mma.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_1_prefill_bench_mma.txt
BATCH=250 KV_LEN=2048 timeout 150 python scripts/decode_shape_bench.py > gpurun_out/r2_1_shape_2b.txt 2>&1; echo "rc=$?"; tail -5 gpurun_out/r2_1_shape_2b.txt
