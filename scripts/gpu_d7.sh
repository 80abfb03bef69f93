#!/bin/bash
mkdir -p gpurun_out
for v in old oldtl default newtl; do
  if [ "$v" = default ]; then unset J3DG_LIB; else export J3DG_LIB=build/variants/libj3dg_$v.so; fi
  echo "[$v]"
  timeout 200 python scripts/perf_cast.py 1184 16 2>&1 | grep -E "^lib=|timeline|rror" | tail -4 | cut -c1-230
done | tee gpurun_out/d7_ab.log
