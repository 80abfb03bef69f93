#!/bin/bash
mkdir -p gpurun_out
{
for f in 8 59 200 500; do
echo -n "[old f=$f] ";  J3DG_LIB=build/variants/libj3dg_old.so timeout 200 python scripts/perf_build.py $f 2>&1 | tail -1
echo -n "[new f=$f] ";  timeout 200 python scripts/perf_build.py $f 2>&1 | tail -1
done
} 2>&1 | tee gpurun_out/e21.log
