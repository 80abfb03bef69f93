#!/bin/bash
mkdir -p gpurun_out
{
echo -n "[prev] "; timeout 200 python scripts/perf_build.py 2>&1 | tail -1
for v in xA xB xC; do
echo -n "[$v] ";  J3DG_LIB=build/variants/libj3dg_$v.so timeout 200 python scripts/perf_build.py 2>&1 | tail -1
done
echo -n "[xA config C] ";  J3DG_LIB=build/variants/libj3dg_xA.so timeout 400 python scripts/perf_build.py 3873 2>&1 | tail -1
J3DG_LIB=build/variants/libj3dg_xA.so timeout 200 python scripts/perf_cast.py 1184 6 2>&1 | tail -1 | cut -c1-300
} 2>&1 | tee gpurun_out/e12.log
J3DG_LIB=build/variants/libj3dg_xA.so timeout 300 ncu --set full --clock-control none --import-source on -k regex:"onesweep_kernel" -s 6 -c 1 -o gpurun_out/prof_e12 -f python scripts/perf_build.py > /dev/null 2>&1
