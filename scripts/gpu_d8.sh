#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
for v in old ws8 ws7; do
  export J3DG_LIB=build/variants/libj3dg_$v.so
  echo -n "[$v] "; LANES=2 timeout 200 python scripts/perf_overlap.py 1184 240 2>&1 | tail -1
done; done | tee gpurun_out/d8_ab.log
for v in ws8 ws7; do J3DG_LIB=build/variants/libj3dg_$v.so LANES=3 timeout 200 python scripts/perf_overlap.py 1184 120 2>&1 | tail -1; done | tee -a gpurun_out/d8_ab.log
J3DG_LIB=build/variants/libj3dg_ws7.so timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not reference and not full_size" 2>&1 | tail -2
