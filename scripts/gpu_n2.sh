#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== dist tests (world 2: peer memory + NCCL staged)"; timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -4 | tee gpurun_out/pytest_dist_2.log
echo "== bench N=2 grid 1024"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --grid 1024 --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_n2_1024.log 2>&1
tail -1 gpurun_out/bench_n2_1024.log | cut -c1-300
