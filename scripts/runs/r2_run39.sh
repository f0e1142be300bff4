#!/bin/bash
# run 39: rehearsal of the round-end checks with the final kernels: pytest -m gpu, smoke(), default N = 1 bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -6 | tee gpurun_out/run39_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | tee gpurun_out/run39_smoke.txt
timeout 900 python bench.py > gpurun_out/run39_bench.json 2> gpurun_out/run39_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/run39_bench.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['config']['step_roofline']['frac_of_hbm_roofline'])
print(d['config']['device_ms_by_class_per_step'])
print(d.get('cpu_baseline',{}).get('value'))
PY
