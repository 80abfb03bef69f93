#!/bin/bash
# final multi-GPU evidence: the driver's invocation at N ranks (N = number of GPUs of the box), plus the C++ group driver
mkdir -p gpurun_out
N=${1:-2}
if [ "$N" = 2 ]; then timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cpp_host or peer" 2>&1 | tail -2; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/f_bench_n$N.json 2> gpurun_out/f_bench_n$N.err; echo "bench n$N rc=$?"; tail -2 gpurun_out/f_bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/f_bench_n$N.json'))
print({k:d[k] for k in ('value','ms_per_step','n_gpus','bvh_broadcast_ms')})
print('single', d['frames_in_flight']['one_frame_at_a_time']['ms_per_frame'], 'e2e', d['e2e']['value'], 'full', d['e2e_full_records']['value'])
print(d['sweep360']['device'], d['sweep360']['e2e']['mrays_s']); print(d.get('exchange'))
c=d['stages'].get('config_c'); print({k:c[k] for k in ('ms_per_frame','mrays_s','sharded_frame_equals_unsharded','bvh_broadcast_ms','mesh_generate_s')})
PY
