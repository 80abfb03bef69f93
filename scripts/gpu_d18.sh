#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in default rf8lk2 rf12 rf16 lk1 rf12lk2; do
  if [ "$v" = default ]; then unset J3DG_LIB; else export J3DG_LIB=build/variants/libj3dg_$v.so; fi
  echo -n "[$v] "; LANES=3 timeout 200 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1
done; done | tee gpurun_out/d19_knobs.log
