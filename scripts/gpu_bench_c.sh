#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload C --f 600 --steps 4 > gpurun_out/bc_small.json 2> gpurun_out/bc_small.err; echo "rc=$?"; tail -3 gpurun_out/bc_small.err
timeout 1200 python bench.py --workload C > gpurun_out/bc_full.json 2> gpurun_out/bc_full.err; echo "rc=$?"; tail -3 gpurun_out/bc_full.err
timeout 600 python bench.py --workload A --steps 60 > gpurun_out/ba.json 2> gpurun_out/ba.err; echo "rc=$?"; tail -3 gpurun_out/ba.err
