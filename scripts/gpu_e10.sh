#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"refit_kernel|collapse_kernel" -s 17 -c 12 -o gpurun_out/prof_e10 -f python scripts/perf_build.py > gpurun_out/e10.log 2>&1
