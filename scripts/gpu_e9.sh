#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 40 -c 80 --csv --log-file gpurun_out/e9_launches.csv python scripts/perf_build.py > /dev/null 2>&1
