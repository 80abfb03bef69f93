#!/bin/bash
# round 2 evidence: default bench (both arms), launch list with DRAM bytes, full ncu captures of the cast / build / splat kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/r2_smi.txt 2>&1
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_steps20.json 2> gpurun_out/r2_bench_steps20.err; echo "bench20 rc=$?"
timeout 900 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-config-c > gpurun_out/r2_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cast_kernel -s 4 -c 1 -o gpurun_out/r2_prof_cast -f python scripts/perf_cast.py 1184 4 > gpurun_out/r2_ncu_cast.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"refit_kernel|collapse_kernel|scatter_kernel|radix_tree_kernel|morton_kernel" -s 30 -c 24 -o gpurun_out/r2_prof_build -f python scripts/perf_build.py > gpurun_out/r2_ncu_build.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"project_kernel|resolve_kernel|seed_kernel|shade_kernel" -s 2 -c 6 -o gpurun_out/r2_prof_splat -f python bench.py --workload P --steps 3 --no-cpu-baseline > gpurun_out/r2_ncu_splat.log 2>&1
ls -la gpurun_out/r2_* | head -30
