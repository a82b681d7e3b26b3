#!/bin/bash
# 8-GPU call: 3LPT 1024^3 (scaling point) and 3LPT 2048^3 (north-star config C5) on one 8xB200 box
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l | tee gpurun_out/gpus_8.txt
echo "== bench N=8 grid 1024"
timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --grid 1024 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n8_1024.log 2>&1
tail -1 gpurun_out/bench_n8_1024.log | cut -c1-300
echo "== bench N=8 grid 2048"
timeout 260 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --grid 2048 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_n8_2048.log 2>&1
tail -3 gpurun_out/bench_n8_2048.log | cut -c1-300
nvidia-smi --query-gpu=memory.used --format=csv,noheader | head -2
