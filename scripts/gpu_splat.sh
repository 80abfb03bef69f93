#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "splat or golden or render_frame or cpp_shim or pick" > gpurun_out/c2_pytest_splat.log 2>&1
echo "exit $?" >> gpurun_out/c2_pytest_splat.log
timeout 300 python bench.py --workload P --steps 5 > gpurun_out/c2_bench_p.json 2> gpurun_out/c2_bench_p.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c2_launches_p.csv python bench.py --workload P --steps 3 --no-cpu-baseline > /dev/null 2>&1
