#!/bin/bash
# env-knob sweep on one variant: J3DG_CONSUMER_BLOCKS x J3DG_LANE_BUDGET
mkdir -p gpurun_out
export J3DG_LIB=build/variants/libj3dg_$1.so
for cb in 0 74 148 296; do for bud in 12 24 40 64; do
  echo -n "consumers=$cb budget=$bud: "
  J3DG_CONSUMER_BLOCKS=$cb J3DG_LANE_BUDGET=$bud timeout 300 python scripts/perf_cast.py 1184 12 2>&1 | grep -E "^lib=|timeline" | tail -2 | sed 's/lib=[^ ]* f=1184 build_ms=[0-9.]* nodes=[0-9]* //'
done; done | tee gpurun_out/knobs.log
