#!/bin/bash
mkdir -p gpurun_out
{
echo -n "[base] ";  J3DG_LIB=build/variants/libj3dg_base.so timeout 200 python scripts/perf_build.py 2>&1 | tail -1
echo -n "[onesweep packed] "; timeout 200 python scripts/perf_build.py 2>&1 | tail -1
echo -n "[onesweep pairs] "; J3DG_SORT=pairs timeout 200 python scripts/perf_build.py 2>&1 | tail -1
echo -n "[lsd] "; J3DG_SORT=lsd timeout 200 python scripts/perf_build.py 2>&1 | tail -1
timeout 200 python scripts/perf_cast.py 1184 6 2>&1 | tail -1 | cut -c1-300
J3DG_SORT=pairs timeout 200 python scripts/perf_cast.py 1184 6 2>&1 | tail -1 | cut -c1-300
timeout 300 python -m pytest tests -x -q -m gpu -k "knn or cube or normals or config_a or rebuild or build" 2>&1 | tail -3
} 2>&1 | tee gpurun_out/e1.log
