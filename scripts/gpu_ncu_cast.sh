#!/bin/bash
# one full ncu capture of the primary cast kernel (config B frame) for the lib given by J3DG_LIB (or the in-tree one)
mkdir -p gpurun_out
name=${1:-cast}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cast_kernel -s 4 -c 1 -o gpurun_out/prof_$name python scripts/perf_cast.py 1184 4 > gpurun_out/ncu_$name.log 2>&1
tail -3 gpurun_out/ncu_$name.log
