#!/bin/bash
# round 2, session 2, call 1: whole GPU suite (incl. the new ingest tests), smoke, bench (both arms), launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/d1_smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > gpurun_out/d1_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/d1_smoke.log
timeout 900 python -m pytest tests/test_ingest.py -q -m gpu > gpurun_out/d1_pytest_ingest.log 2>&1; echo "ingest rc=$?"; tail -15 gpurun_out/d1_pytest_ingest.log
timeout 1500 python -m pytest tests -x -q -m gpu --deselect tests/test_ingest.py > gpurun_out/d1_pytest_all.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/d1_pytest_all.log
timeout 600 python bench.py > gpurun_out/d1_bench.json 2> gpurun_out/d1_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/d1_bench.json
timeout 300 python bench.py --impl reference --steps 6 --warmup 1 > gpurun_out/d1_bench_ref.json 2> gpurun_out/d1_bench_ref.err; cut -c1-300 gpurun_out/d1_bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/d1_launches.csv python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/d1_bench_under_ncu.log 2>&1
echo done
