#!/bin/bash
mkdir -p gpurun_out
{
echo -n "[prev] "; timeout 200 python scripts/perf_build.py 2>&1 | tail -1
for v in nA; do
echo -n "[$v] ";  J3DG_LIB=build/variants/libj3dg_$v.so timeout 200 python scripts/perf_build.py 2>&1 | tail -1
done
J3DG_LIB=build/variants/libj3dg_nA.so timeout 200 python scripts/perf_cast.py 1184 6 2>&1 | tail -1 | cut -c1-300
J3DG_LIB=build/variants/libj3dg_nA.so timeout 300 python -m pytest tests -x -q -m gpu -k "cube or config_a or rebuild or build or degenerate or knn or ply or random_transformed" 2>&1 | tail -3
echo -n "[nA config C] "; J3DG_LIB=build/variants/libj3dg_nA.so timeout 400 python scripts/perf_build.py 3873 2>&1 | tail -1
} 2>&1 | tee gpurun_out/e22.log
