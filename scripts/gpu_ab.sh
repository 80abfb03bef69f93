#!/bin/bash
# A/B of cast-kernel variants on the GPU: (optionally) correctness first, then perf per variant.
# usage: [TESTS=1] bash scripts/gpu_ab.sh variant...   (variants are build/variants/libj3dg_<name>.so; "default" = in-tree)
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" > gpurun_out/pytest_ab.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_ab.log
fi
for v in default "$@"; do
  if [ "$v" = default ]; then unset J3DG_LIB; else export J3DG_LIB=build/variants/libj3dg_$v.so; fi
  timeout 300 python scripts/perf_cast.py 2>&1 | grep -E "^lib=|timeline" | tail -4
done | tee gpurun_out/ab.log
