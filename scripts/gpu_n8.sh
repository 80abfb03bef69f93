#!/bin/bash
# One 8-GPU call: scaling bench at N=8 and N=4 (orbit sweep sharded frame-wise, gather overlapped) and BASELINE
# configs[2] (300 M triangles, 4K + shadows, screen bands over 8 GPUs).  Outputs in gpurun_out/.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
  timeout 300 $TR --nproc-per-node $n --master-port $((29500 + n)) bench.py --gpus $n --steps 120 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  echo "bench n=$n rc=$?"; tail -1 gpurun_out/bench_n$n.json | cut -c1-600
done
timeout 600 $TR --nproc-per-node 8 --master-port 29520 scripts/config_c_tiles.py --f 3873 --frames 8 > gpurun_out/config_c_n8.json 2> gpurun_out/config_c_n8.err
echo "config C n=8 rc=$?"; tail -1 gpurun_out/config_c_n8.json
