#!/bin/bash
# 8-GPU orbit sweep: frame exchange through NVLink peer memory vs NCCL gather (plain, and with copy-engine P2P).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29508 bench.py --gpus 8 --steps 120 --warmup 5 --exchange peer > gpurun_out/bench_n8_peer.json 2> gpurun_out/bench_n8_peer.err
echo "peer rc=$?"; tail -1 gpurun_out/bench_n8_peer.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('exchange'), d['e2e']['value'])"; tail -3 gpurun_out/bench_n8_peer.err
timeout 300 $TR --nproc-per-node 8 --master-port 29509 bench.py --gpus 8 --steps 120 --warmup 5 --exchange nccl > gpurun_out/bench_n8_nccl.json 2> gpurun_out/bench_n8_nccl.err
echo "nccl rc=$?"; tail -1 gpurun_out/bench_n8_nccl.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('exchange'))"
timeout 300 $TR --nproc-per-node 2 --master-port 29511 bench.py --gpus 2 --steps 120 --warmup 5 > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err
echo "n2 peer rc=$?"; tail -1 gpurun_out/bench_n2_peer.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d.get('exchange'))"
