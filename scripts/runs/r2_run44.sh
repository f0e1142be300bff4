#!/bin/bash
# run 44: final validation of the committed kernels: pytest -m gpu, smoke(), default bench; then the ncu launch list of one
# whole decode step (gpu__time_duration only, clocks as found -- per-launch times are cold and serialised, shares are what count)
# and one --set full capture of the transposed W4A16 kernel (gate_up shape of 70B at TP = 8)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/run44_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/run44_smoke.txt
timeout 900 python bench.py --no-alt > gpurun_out/run44_bench.json 2> gpurun_out/run44_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/run44_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['config']['step_roofline']['frac_of_hbm_roofline'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/run44_launches.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu --no-alt --kv-budget-tokens 526000 > gpurun_out/run44_ncu_bench.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/run44_launches.csv
FUSED_ONLY=1 SHAPES=gate_up timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_w4t_kernel -s 6 -c 1 \
  -o gpurun_out/run44_w4t_gateup -f python scripts/gemm_w4_bench.py > gpurun_out/run44_ncu_w4t.log 2>&1; echo "w4t capture rc=$?"
ls -la gpurun_out/run44_*
