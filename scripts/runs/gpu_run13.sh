#!/bin/bash
# round 1, run 13 (last GPU minutes of the round): validate the ONNX model-slice loader on the device, rehearse the
# driver's round-end sequence (smoke, default bench), refresh the launch list, then as much of the GPU suite as fits.
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k "pmx_onnx" > gpurun_out/pytest13_onnx.log 2>&1; echo "rc=$?" | tee -a gpurun_out/pytest13_onnx.log; tail -12 gpurun_out/pytest13_onnx.log | cut -c1-300
timeout 120 python __graft_entry__.py smoke > gpurun_out/smoke13.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke13.log | cut -c1-300
timeout 300 python bench.py > gpurun_out/bench13.json 2> gpurun_out/bench13.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/bench13.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches13.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu13.log 2>&1; echo "ncu rc=$?"; wc -l gpurun_out/launches13.csv
timeout 600 python -m pytest tests -x -q -m gpu --deselect tests/test_host_cpp.py::test_reference_generator_from_pmx_onnx_export > gpurun_out/pytest13_all.log 2>&1; echo "rc=$?" | tee -a gpurun_out/pytest13_all.log; tail -6 gpurun_out/pytest13_all.log | cut -c1-300
