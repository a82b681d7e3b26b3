#!/bin/bash
# N-GPU call: slab parity tests on min(N,..) ranks, then the 1024^3 bench at N GPUs (and smaller grids).
set -u
N=${NG:-2}
mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/gpus_$N.txt
echo "== dist tests"; timeout 900 python -m pytest tests/test_gpu_dist.py -x -q 2>&1 | tail -8 | tee gpurun_out/pytest_dist_$N.log
for G in ${GRIDS:-512 1024}; do
echo "== bench N=$N grid $G"
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --grid $G --steps 3 --warmup 3 --no-cpu 2>&1 | tail -2 | tee gpurun_out/bench_n${N}_$G.log | cut -c1-250
done
