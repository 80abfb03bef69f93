#!/bin/bash
mkdir -p gpurun_out
( for l in 2 3 4; do LANES=$l timeout 300 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1; done
for b in 16 20 32 40 56 80; do J3DG_LANE_BUDGET=$b LANES=2 timeout 300 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1; done ) | tee gpurun_out/d5_overlap.log
