#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/dist_worker.py --grid 64 --lattice bcc > gpurun_out/dist_dbg.log 2>&1
tail -40 gpurun_out/dist_dbg.log
