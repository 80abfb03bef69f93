#!/bin/bash
# 2-GPU call: the C++ group driver, the peer-frames tests, bench at N = 2 (workloads B and C)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_smi.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "group_abi or peer_frames" > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/n2_pytest.log
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/n2_bench_b.json 2> gpurun_out/n2_bench_b.err; echo "bench B rc=$?"; tail -5 gpurun_out/n2_bench_b.err
P=$((20000 + RANDOM % 20000))
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --workload C --f 1184 --steps 8 > gpurun_out/n2_bench_c.json 2> gpurun_out/n2_bench_c.err; echo "bench C rc=$?"; tail -5 gpurun_out/n2_bench_c.err
