#!/bin/bash
# run 49: W4A16 shapes with more work items than SMs (persistent loop, accumulator reuse)
timeout 80 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "test_gemm_w4a16_fused and (20480 or 12800)" -x 2>&1 | tail -8
