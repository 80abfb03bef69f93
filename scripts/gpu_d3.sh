#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "two_contexts or cpp_host or peer or pipelined" > gpurun_out/d3_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d3_pytest.log
timeout 900 python bench.py --no-splat > gpurun_out/d3_bench.json 2> gpurun_out/d3_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/d3_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d3_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','cast_ms','shade_ms')})
print(d['frames_in_flight']['one_frame_at_a_time'])
print(d['e2e']); print(d['e2e_full_records']); print(d['sweep360']); print(d['roofline']['frac'], d['roofline']['achieved_frames_in_flight']); print(d.get('parity'))
PY
