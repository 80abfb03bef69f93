#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/e13_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/e13_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/e13_bench.json 2> gpurun_out/e13_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/e13_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/e13_bench.json'))
print({k:d[k] for k in ('value','steps','ms_per_step','cast_ms','shade_ms','bvh_build_ms','bvh_nodes')})
print('  single', d['frames_in_flight']['one_frame_at_a_time']['ms_per_frame'], 'e2e', d['e2e']['value'], 'mesh_create', d['e2e'].get('mesh_create_ms'))
print('  roof', d['roofline']['frac'], 'parity', d['parity']['pass'], d['parity']['id_mismatch'], 'build', d['stages']['build'])
print('  C', {k:v for k,v in d['stages']['config_c'].items() if k!='workload'})
print('  cpu', d['cpu_baseline'])
PY
