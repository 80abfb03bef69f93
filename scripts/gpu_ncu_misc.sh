#!/bin/bash
# Full ncu captures of the kernels around the cast kernel (one launch each): shade, resolve, refit, radix tree, collapse (largest level), all-hits.
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"shade_kernel|resolve_kernel" -s 6 -c 2 -o gpurun_out/prof_shade_resolve -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_misc1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"refit_kernel|radix_tree_kernel" -s 2 -c 2 -o gpurun_out/prof_refit_radix -f python scripts/perf_build.py > gpurun_out/ncu_misc2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"allhits_kernel" -s 3 -c 1 -o gpurun_out/prof_allhits -f python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_misc3.log 2>&1
ls -la gpurun_out/*.ncu-rep
