#!/bin/bash
# N-GPU runs: bench.py orbit sweep + configs[2] screen sharding.  usage: bash scripts/gpu_multi.sh N [f_for_config_c]
N=$1; F=${2:-1184}
mkdir -p gpurun_out
run() { if [ "$N" = 1 ]; then python "$@"; else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 "$@"; fi; }
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
run bench.py --gpus $N --steps 120 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"; cat gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
run scripts/config_c_tiles.py --f $F --frames 8 > gpurun_out/config_c_n$N.json 2> gpurun_out/config_c_n$N.err; echo "config_c rc=$?"; cat gpurun_out/config_c_n$N.json; tail -3 gpurun_out/config_c_n$N.err
