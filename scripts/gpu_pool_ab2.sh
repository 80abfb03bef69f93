#!/bin/bash
# perf A/B: spec = variant[:budget[:consumer_blocks]]
mkdir -p gpurun_out
for spec in "$@"; do
  IFS=: read v budget cons <<< "$spec"
  if [ "$v" = default ]; then unset J3DG_LIB; else export J3DG_LIB=build/variants/libj3dg_$v.so; fi
  if [ -n "$budget" ]; then export J3DG_LANE_BUDGET=$budget; else unset J3DG_LANE_BUDGET; fi
  if [ -n "$cons" ]; then export J3DG_CONSUMER_BLOCKS=$cons; else unset J3DG_CONSUMER_BLOCKS; fi
  echo -n "[$spec] "
  timeout 200 python scripts/perf_cast.py 2>&1 | grep -E "^lib=|timeline|rror" | tail -2 | cut -c1-230
done | tee gpurun_out/pool_ab2.log
