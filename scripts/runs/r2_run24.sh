#!/bin/bash
# round 2, run 24 (1 x B200): sampler with 16-byte loads -- parity tests, then its time inside the step (ncu, one launch)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_ops_gpu.py -m gpu -x -q -k "sampler or penalty" > gpurun_out/r2_24_sampler.log 2>&1; echo "rc=$?" >> gpurun_out/r2_24_sampler.log; tail -4 gpurun_out/r2_24_sampler.log | cut -c1-300
timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -x -q -k "w8a8_paged or config1 or errors" > gpurun_out/r2_24_engine.log 2>&1; echo "rc=$?" >> gpurun_out/r2_24_engine.log; tail -3 gpurun_out/r2_24_engine.log | cut -c1-300
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:sample_kernel -c 3 --csv --log-file gpurun_out/r2_24_sampler_ncu.csv python bench.py --layers 2 --kv-len 512 --kv-budget-tokens 524288 --steps 2 --warmup 3 --no-cpu --no-alt > gpurun_out/r2_24_ncu.log 2>&1; grep sample_kernel gpurun_out/r2_24_sampler_ncu.csv | cut -d, -f5,12- | head -6
