#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_cpp.py tests/test_tp_gpu.py -q -m gpu -k "tensor_parallel" > gpurun_out/pytest_tp.log 2>&1; echo "tp rc=$?" | tee -a gpurun_out/pytest_tp.log
tail -30 gpurun_out/pytest_tp.log | cut -c1-250
timeout 400 python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py -x -q -m gpu -k "w4a16 or config4 or config3 or prefill" > gpurun_out/pytest_w4.log 2>&1; echo "w4 rc=$?" | tee -a gpurun_out/pytest_w4.log
tail -15 gpurun_out/pytest_w4.log | cut -c1-250
for sms in 148 96 64 32; do B2LLM_GEMM_SMS=$sms timeout 120 python scripts/overlap_probe.py; done > gpurun_out/overlap.log 2>&1
cat gpurun_out/overlap.log
timeout 600 python scripts/steady_state.py > gpurun_out/steady.log 2>&1; tail -45 gpurun_out/steady.log
