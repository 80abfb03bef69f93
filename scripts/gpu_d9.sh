#!/bin/bash
mkdir -p gpurun_out
( for v in ws7 ws6; do export J3DG_LIB=build/variants/libj3dg_$v.so; echo -n "[$v] "; LANES=3 timeout 200 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1; done
export J3DG_LIB=build/variants/libj3dg_ws7.so
for b in 20 28 32 40; do echo -n "[ws7] "; J3DG_LANE_BUDGET=$b LANES=3 timeout 200 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1; done
echo -n "[ws7 4 lanes] "; LANES=4 timeout 200 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1 ) | tee gpurun_out/d9_sweep.log
