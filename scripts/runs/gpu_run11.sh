#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_host_cpp.py -x -q -m gpu -k "prefix_cache or penalty" > gpurun_out/pytest11.log 2>&1; echo "rc=$?" | tee -a gpurun_out/pytest11.log; tail -15 gpurun_out/pytest11.log | cut -c1-250
timeout 60 python scripts/hbm_sustained.py > gpurun_out/hbm_sustained.log 2>&1; cat gpurun_out/hbm_sustained.log
# one whole decode step (2 layers, B=1024, kv_len 512) under --set full; KV allocation = exactly the tokens in use, because ncu
# saves / restores all device memory on every replay pass (a 170 GB cache makes each pass take seconds)
timeout 500 ncu --set full --clock-control none --import-source on -k regex:"attn_|gemm_tc|rmsnorm|quant_rows|rope_kv|sample_kernel|embedding|gather_rows" -s 75 -c 25 -o gpurun_out/prof_step11 -f python bench.py --layers 2 --kv-len 512 --kv-budget-tokens 524288 --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_step11.log 2>&1
tail -2 gpurun_out/ncu_step11.log
