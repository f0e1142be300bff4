#!/bin/bash
# 1-GPU session: full GPU test suite, bench (+ CPU baseline), reference arm, launch list, ncu --set full of every step kernel
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench7.json 2> gpurun_out/bench7.err; echo "bench rc=$?"
cat gpurun_out/bench7.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench7_ref.json 2>> gpurun_out/bench7.err; cat gpurun_out/bench7_ref.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches7.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_|gemm_tc|rmsnorm|quant_rows|rope_kv|sample_kernel|embedding|gather_rows" -s 75 -c 25 -o gpurun_out/prof_step7 -f python bench.py --layers 2 --kv-len 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_step7.log 2>&1
tail -3 gpurun_out/ncu_step7.log
timeout 120 python scripts/prefill_bench.py > gpurun_out/prefill_bench.log 2>&1; cat gpurun_out/prefill_bench.log
timeout 200 python scripts/gemm_bench.py 8192 > gpurun_out/gemm_bench_8192.log 2>&1; cat gpurun_out/gemm_bench_8192.log
ls -la gpurun_out | head -40
