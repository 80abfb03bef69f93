#!/bin/bash
# batch 1: warp handover / next-child prefetch variants — parity of the combined variant, then perf of each
mkdir -p gpurun_out
J3DG_LIB=build/variants/libj3dg_whpf.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" > gpurun_out/pytest_whpf.log 2>&1; echo "pytest(whpf) rc=$?"; tail -3 gpurun_out/pytest_whpf.log
bash scripts/gpu_ab.sh wh pf whpf
for b in 16 20 32; do
  J3DG_LANE_BUDGET=$b J3DG_LIB=build/variants/libj3dg_whpf.so timeout 300 python scripts/perf_cast.py 2>&1 | grep -E "^lib=" | sed "s/^/budget=$b /"
done | tee -a gpurun_out/ab.log
