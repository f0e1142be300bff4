#!/bin/bash
# round 2, run 18 (2 x B200): TP tests incl. the peer-mapping-failure fallback; the reference's single-process TP host
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tp_gpu.py -x -q -m gpu > gpurun_out/r2_18_tp.log 2>&1; echo "rc=$?" >> gpurun_out/r2_18_tp.log; tail -6 gpurun_out/r2_18_tp.log | cut -c1-400
timeout 400 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k "tensor_parallel_2" > gpurun_out/r2_18_tp_host.log 2>&1; echo "rc=$?" >> gpurun_out/r2_18_tp_host.log; tail -4 gpurun_out/r2_18_tp_host.log | cut -c1-300
