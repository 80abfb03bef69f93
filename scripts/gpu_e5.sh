#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"onesweep_kernel|radix_tree" -s 6 -c 3 -o gpurun_out/prof_e5b -f python scripts/perf_build.py > gpurun_out/e5.log 2>&1
ls -la gpurun_out/*.ncu-rep
