#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/d10_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/d10_smoke.log | cut -c1-200
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/d10_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/d10_pytest.log
timeout 900 python bench.py > gpurun_out/d10_bench.json 2> gpurun_out/d10_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/d10_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d10_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','cast_ms','shade_ms')})
print(d['frames_in_flight']['one_frame_at_a_time'])
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'full', d['e2e_full_records']['value'], d['e2e_full_records']['full_copy']['value'], d['e2e_full_records']['sync_render_frame_ms_per_step'])
print(d['sweep360']); print(d['roofline']['frac'], d['roofline']['achieved_frames_in_flight']); print(d.get('parity'))
print(d['stages'].get('splat',{}).get('ms'))
PY
