#!/bin/bash
mkdir -p gpurun_out
( echo "[default]"; timeout 200 python scripts/perf_cast.py 1184 8 2>&1 | tail -1 | cut -c1-260
echo "[fused]"; J3DG_BUILD_FUSED=1 timeout 200 python scripts/perf_cast.py 1184 8 2>&1 | tail -1 | cut -c1-260
echo "[default build]"; timeout 200 python scripts/perf_build.py 2>&1 | tail -1
echo "[fused build]"; J3DG_BUILD_FUSED=1 timeout 200 python scripts/perf_build.py 2>&1 | tail -1
echo "[fused build f=3873]"; J3DG_BUILD_FUSED=1 timeout 300 python scripts/perf_build.py 3873 2>&1 | tail -1
echo "[default build f=3873]"; timeout 300 python scripts/perf_build.py 3873 2>&1 | tail -1 ) | tee gpurun_out/d25_fused.log
J3DG_BUILD_FUSED=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "not reference and not cpp" 2>&1 | tail -2
