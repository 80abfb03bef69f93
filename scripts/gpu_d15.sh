#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 40 --warmup 5 --no-config-c --no-splat > gpurun_out/d15_bench.json 2> gpurun_out/d15_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/d15_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/d15_bench.json'))
print(json.dumps(d['stages']['ingest'], indent=1))
print(d['value'])
PY
