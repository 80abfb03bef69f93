#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "cpp_host or peer" > gpurun_out/d4_pytest_n2.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d4_pytest_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 120 --warmup 6 > gpurun_out/d4_bench_n2.json 2> gpurun_out/d4_bench_n2.err; echo "bench n2 rc=$?"; tail -3 gpurun_out/d4_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d4_bench_n2.json'))
print({k:d[k] for k in ('value','ms_per_step','cast_ms','shade_ms','n_gpus')})
print(d['frames_in_flight']['one_frame_at_a_time'])
print(d['e2e']); print(d['sweep360']); print(d.get('exchange'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload C --f 1184 --steps 10 --warmup 3 > gpurun_out/d4_bench_c_n2.json 2> gpurun_out/d4_bench_c_n2.err; echo "bench C n2 rc=$?"; cut -c1-400 gpurun_out/d4_bench_c_n2.json
