#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -3 > gpurun_out/r2_n8_smi.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; echo "bench n8 rc=$?"; tail -3 gpurun_out/r2_bench_n8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print({k:d[k] for k in ('value','ms_per_step','cast_ms','shade_ms','n_gpus','bvh_broadcast_ms')})
print(d['frames_in_flight']['one_frame_at_a_time'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'full', d['e2e_full_records']['value'])
print(d['sweep360']); print(d.get('exchange')); print(d['stages'].get('config_c'))
PY
