#!/bin/bash
# run 48: the whole GPU test suite on the final commit (planner change included)
mkdir -p gpurun_out
timeout 260 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/run48_pytest.txt
