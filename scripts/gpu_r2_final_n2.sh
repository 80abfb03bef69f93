#!/bin/bash
# round-2 final 2-GPU check with the rebuilt builder: the world-2 C++ group driver test and the driver's bench invocation at N = 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "group_abi" > gpurun_out/g2_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g2_pytest.log
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/g2_bench.json 2> gpurun_out/g2_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/g2_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/g2_bench.json'))
print({k:d[k] for k in ('value','n_gpus','steps','ms_per_step','bvh_build_ms')}, 'e2e', d['e2e']['value'])
print('sweep', d['sweep360']['device'], 'C', d['stages']['config_c']['ms_per_frame'], d['stages']['config_c']['sharded_frame_equals_unsharded'], d['stages']['config_c'].get('bvh_broadcast_ms'))
PY
