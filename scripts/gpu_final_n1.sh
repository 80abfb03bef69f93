#!/bin/bash
# final single-GPU evidence of the round: smoke, whole GPU suite, bench (driver's invocation, default, reference arm)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/f_smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log | cut -c1-160
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/f_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/f_bench_steps20.json 2> gpurun_out/f_bench_steps20.err; echo "bench20 rc=$?"
timeout 900 python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
python - <<'PY'
import json
for f in ('gpurun_out/f_bench_steps20.json','gpurun_out/f_bench.json'):
    d=json.load(open(f))
    print(f, {k:d[k] for k in ('value','steps','ms_per_step','cast_ms','shade_ms','bvh_build_ms','bvh_nodes')})
    print('  single', d['frames_in_flight']['one_frame_at_a_time']['ms_per_frame'], 'e2e', d['e2e']['value'], 'full', d['e2e_full_records']['value'], 'sync', d['e2e_full_records']['sync_render_frame_ms_per_step'], d['e2e_full_records'].get('sync_render_frame_dirty_rect_ms_per_step'))
    print('  roof', d['roofline']['frac'], d['roofline']['achieved_frames_in_flight'], 'parity', d['parity']['pass'], d['parity']['id_mismatch'])
    print('  C', d['stages']['config_c']['ms_per_frame'], d['stages']['config_c']['sharded_frame_equals_unsharded'], 'splat', d['stages']['splat']['ms'], 'ply', d['stages']['ingest']['ply']['decode_kernels_ms'], 'knn', d['stages']['ingest']['normals']['knn_fit_ms'])
PY
