#!/bin/bash
mkdir -p gpurun_out
for spl in 1 2 1 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2953$spl bench.py --gpus 8 --steps 60 --warmup 6 --slots-per-lane $spl --no-config-c --no-splat --no-ingest --no-cpu-baseline > gpurun_out/d17_n8_spl$spl.json 2> gpurun_out/d17_n8_spl$spl.err
python - <<PY
import json
d=json.load(open('gpurun_out/d17_n8_spl$spl.json'))
print('slots/lane $spl', round(d['value']), round(d['ms_per_step'],4), 'sweep360', round(d['sweep360']['device']['mrays_s']), 'e2e', round(d['e2e']['value']), d.get('exchange'))
PY
done 2>&1 | tee gpurun_out/d17_slots.log
