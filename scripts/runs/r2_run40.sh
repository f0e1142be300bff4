#!/bin/bash
# run 40 / 47 (8 GPUs; RUN=47 for the rerun): the 70B GQA W4A16 TP = 8 step (BASELINE configs[3], literal shape).
# Run 40: with the transposed W4A16 kernel.  Run 47: + the single-wave decode attention plan.
# The replica leg is shortened (kv_len 64) and the TP leg filtered to the 70B shape to keep the 8x-charged call short.
mkdir -p gpurun_out
export RUN=${RUN:-40}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 8 --steps 10 --warmup 3 --kv-len 64 --no-cpu --tp-filter 70B > gpurun_out/run${RUN}_n8_70b.json 2> gpurun_out/run${RUN}_n8_70b.err
echo "bench rc=$?"
python - <<'PY'
import json, os
d=json.loads(open(f"gpurun_out/run{os.environ['RUN']}_n8_70b.json").read().strip().splitlines()[-1])
tp=d['config']['tp']
print(json.dumps(tp['parity_gate']))
for r in tp['runs']:
    print({k:r.get(k) for k in ('model','tp','batch','kv_len','ms_per_step','tokens_per_s','e2e_tokens_per_s','error')})
    print(r.get('device_ms_by_class_per_step'))
    print(r.get('per_gpu_roofline'), r.get('fused_join_us_per_call'))
PY
tail -3 gpurun_out/run${RUN}_n8_70b.err
