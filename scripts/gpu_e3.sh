#!/bin/bash
mkdir -p gpurun_out
{
echo -n "[prev] "; timeout 200 python scripts/perf_build.py 2>&1 | tail -1
echo -n "[rA] ";  J3DG_LIB=build/variants/libj3dg_rA.so timeout 200 python scripts/perf_build.py 2>&1 | tail -1
J3DG_LIB=build/variants/libj3dg_rA.so timeout 200 python scripts/perf_cast.py 1184 6 2>&1 | tail -1 | cut -c1-300
J3DG_LIB=build/variants/libj3dg_rA.so timeout 300 python -m pytest tests -x -q -m gpu -k "cube or config_a or rebuild or build or tuning or ragged or empty" 2>&1 | tail -3
J3DG_LIB=build/variants/libj3dg_rA.so timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/e3_launches.csv python scripts/perf_build.py > /dev/null 2>&1
} 2>&1 | tee gpurun_out/e3.log
