#!/bin/bash
# run 26: first device run of tests/test_golden_gpu.py (prints measured cache/logit differences), then the engine tests
mkdir -p gpurun_out
python -m pytest tests/test_golden_gpu.py -q -m gpu -s -x > gpurun_out/run26_golden.log 2>&1
echo "golden rc=$?" >> gpurun_out/run26_golden.log
grep -E "GOLDEN|passed|failed|rc=" gpurun_out/run26_golden.log
tail -30 gpurun_out/run26_golden.log
