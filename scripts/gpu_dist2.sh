#!/bin/bash
# 2-GPU call: slab parity tests (peer-memory fused + NCCL staged), then short benches at N=2.
set -u
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus.txt
nvidia-smi topo -m 2>&1 | head -12 | tee -a gpurun_out/gpus.txt
echo "== dist tests"; timeout 600 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_dist.log
for G in ${GRIDS:-256 512}; do
echo "== bench N=2 grid $G"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --grid $G --steps 3 --warmup 3 --no-cpu 2>&1 | tail -4 | tee gpurun_out/bench_n2_$G.log
done
echo "== bench N=2 grid 512 NCCL-staged exchange"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --grid 512 --steps 3 --warmup 3 --no-cpu --no-e2e --staged-exchange 2>&1 | tail -4 | tee gpurun_out/bench_n2_512_staged.log
