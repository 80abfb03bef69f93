#!/bin/bash
# ncu launch list (time + DRAM bytes) of the BVH build alone: scripts/perf_build.py, config B
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 33 -c 66 --csv --log-file gpurun_out/build_launches.csv python scripts/perf_build.py > /dev/null 2>&1
