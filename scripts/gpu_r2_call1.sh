#!/bin/bash
# round 2, GPU call 1: the new reference-parity tests, the whole GPU suite, the new bench line, splat / build profiles
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc >> gpurun_out/c1_smi.txt; free -g | head -2 >> gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "reference or peer_frames or cpp_shim" > gpurun_out/c1_pytest_new.log 2>&1
echo "exit $?" >> gpurun_out/c1_pytest_new.log
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_gpu_parity.py::test_config_b_1080p_matches_reference > gpurun_out/c1_pytest_all.log 2>&1
echo "exit $?" >> gpurun_out/c1_pytest_all.log
timeout 600 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "exit $?" >> gpurun_out/c1_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"project_kernel|seed_kernel|resolve_kernel" -s 4 -c 4 -o gpurun_out/c1_prof_splat python bench.py --workload P --steps 3 --no-cpu-baseline > gpurun_out/c1_ncu_splat.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/c1_launches_p.csv python bench.py --workload P --steps 3 --no-cpu-baseline > /dev/null 2>&1
