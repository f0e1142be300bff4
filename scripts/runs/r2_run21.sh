#!/bin/bash
# round 2, run 21 (1 x B200): the two host-tool tests still gated since round 1 (prefix-cache benchmark tool, trained sentencepiece
# model through offline_inference) + the new W4A16 test shape
mkdir -p gpurun_out
B2LLM_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_host_cpp.py -m gpu -x -q -k "prefix_cache_benchmark or trained_sentencepiece" > gpurun_out/r2_21_host_gated.log 2>&1; echo "rc=$?" >> gpurun_out/r2_21_host_gated.log; tail -15 gpurun_out/r2_21_host_gated.log | cut -c1-400
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "w4a16" > gpurun_out/r2_21_w4.log 2>&1; echo "rc=$?" >> gpurun_out/r2_21_w4.log; tail -4 gpurun_out/r2_21_w4.log | cut -c1-300
