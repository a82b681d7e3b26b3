#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench N=4 grid 1024"
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --grid 1024 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n4_1024.log 2>&1
tail -1 gpurun_out/bench_n4_1024.log | cut -c1-300
