#!/bin/bash
# round-2 final single-GPU evidence with the rebuilt BVH builder: smoke, whole GPU suite, bench (driver's invocation, default, reference arm),
# launch list with DRAM bytes, full ncu capture of one build (every kernel of the new builder)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/g_smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/g_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/g_smoke.log | cut -c1-200
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/g_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/g_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/g_bench_steps20.json 2> gpurun_out/g_bench_steps20.err; echo "bench20 rc=$?"
timeout 900 python bench.py > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/g_bench_ref.json 2> gpurun_out/g_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/g_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-config-c > gpurun_out/g_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"morton_kernel|digit_histograms_kernel|onesweep_kernel|tree_fit_kernel|upper_tree_kernel|climb_kernel|collapse_kernel" -s 20 -c 22 -o gpurun_out/g_prof_build -f python scripts/perf_build.py > gpurun_out/g_ncu_build.log 2>&1
echo "[config C build] $(timeout 400 python scripts/perf_build.py 3873 2>&1 | tail -1)" | tee gpurun_out/g_build_c.log
python - <<'PY'
import json
for f in ('gpurun_out/g_bench_steps20.json','gpurun_out/g_bench.json'):
    d=json.load(open(f))
    print(f, {k:d[k] for k in ('value','steps','ms_per_step','cast_ms','shade_ms','bvh_build_ms','bvh_nodes')})
    print('  single', d['frames_in_flight']['one_frame_at_a_time']['ms_per_frame'], 'e2e', d['e2e']['value'], 'full', d['e2e_full_records']['value'], 'mesh_create', d['e2e'].get('mesh_create_ms'))
    print('  roof', d['roofline']['frac'], d['roofline']['achieved_frames_in_flight'], 'parity', d['parity']['pass'], d['parity']['id_mismatch'], 'build', d['stages']['build'])
    print('  C', d['stages']['config_c']['ms_per_frame'], d['stages']['config_c']['bvh_build_ms'], d['stages']['config_c']['sharded_frame_equals_unsharded'], 'splat', d['stages']['splat']['ms'], 'ply', d['stages']['ingest']['ply'].get('decode_kernels_ms'), d['stages']['ingest']['ply'].get('mesh_create_from_ply_ms'))
print(open('gpurun_out/g_bench_ref.json').read()[:600])
PY
