#!/bin/bash
# round 2, run 25 (1 x B200): ncu evidence with the final kernels -- one decode step (--set full), the tcgen05 prefill kernel
# (hi + lo P), the W4A16 GEMM at the 70B / TP 8 shapes
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__cycles_elapsed.avg.per_second"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_decode|gemm_tc2|row_quant|rmsnorm_quant|quant_rows|rope_kv|sample_kernel" -s 60 -c 14 -o gpurun_out/r2_25_step -f python bench.py --layers 2 --kv-len 512 --kv-budget-tokens 524288 --steps 1 --warmup 3 --no-cpu --no-alt > gpurun_out/r2_25_ncu_step.log 2>&1; echo "step capture rc=$?"
ncu -i gpurun_out/r2_25_step.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2_25_step_raw.csv 2>&1; wc -l gpurun_out/r2_25_step_raw.csv
IMPL=6 SEQS=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc -c 1 -o gpurun_out/r2_25_prefill_tc -f python scripts/prefill_bench.py > gpurun_out/r2_25_ncu_prefill.log 2>&1; echo "prefill capture rc=$?"
ncu -i gpurun_out/r2_25_prefill_tc.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2_25_prefill_raw.csv 2>&1
ncu -i gpurun_out/r2_25_prefill_tc.ncu-rep --page details > gpurun_out/r2_25_prefill_tc_details.txt 2>&1; grep -E "Duration|TC is|Issue Slots Busy|Registers Per|Achieved Occupancy" gpurun_out/r2_25_prefill_tc_details.txt | head
timeout 300 ncu --set full --clock-control none -k regex:"gemm_w4_kernel|w4_splitk" -s 8 -c 6 -o gpurun_out/r2_25_gemm_w4 -f python scripts/gemm_w4_bench.py > gpurun_out/r2_25_ncu_w4.log 2>&1; echo "w4 capture rc=$?"
ncu -i gpurun_out/r2_25_gemm_w4.ncu-rep --page raw --csv --metrics $M > gpurun_out/r2_25_gemm_w4_raw.csv 2>&1; cut -c1-400 gpurun_out/r2_25_gemm_w4_raw.csv | tail -8
rm -f gpurun_out/r2_25_gemm_w4.ncu-rep
