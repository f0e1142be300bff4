#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log | cut -c1-250
timeout 100 python scripts/gemm_w4_bench.py 256 > gpurun_out/gemm_w4_bench.log 2>&1; cat gpurun_out/gemm_w4_bench.log
# every step kernel under --set full; the CTA-pair GEMM does not survive ncu's multi-pass replay (run 7), so this capture
# uses the single-CTA tcgen05 kernel (B2LLM_GEMM_2CTA=0) and the pair kernel gets a light single-metric pass below
B2LLM_GEMM_2CTA=0 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_|gemm_tc|rmsnorm|quant_rows|rope_kv|sample_kernel|embedding|gather_rows" -s 75 -c 25 -o gpurun_out/prof_step10 -f python bench.py --layers 2 --kv-len 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_step10.log 2>&1
tail -2 gpurun_out/ncu_step10.log
timeout 150 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gemm_tc2" -s 12 -c 4 --csv --log-file gpurun_out/pair_light.csv python bench.py --layers 2 --kv-len 512 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_pair.log 2>&1; tail -2 gpurun_out/ncu_pair.log; cat gpurun_out/pair_light.csv | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench10.json 2> gpurun_out/bench10.err; echo "bench rc=$?"
cat gpurun_out/bench10.json
