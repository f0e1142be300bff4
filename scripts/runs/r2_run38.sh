#!/bin/bash
# run 38 (2 GPUs): tensor-parallel tests (W8A8 / fp16 / W4A16 x fused / nccl join) with the transposed W4A16 kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_tp_gpu.py -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/run38_tp_tests.txt
