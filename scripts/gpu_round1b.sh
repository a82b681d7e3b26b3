#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== diag"; timeout 300 python scripts/gpu_diag.py 2>&1 | tail -20 | tee gpurun_out/diag.log
echo "== bench 1024"; timeout 1200 python bench.py --steps 3 --warmup 3 --cpu-grid 128 2>&1 | tail -3 | tee gpurun_out/bench_1024.log
SKIP=823 COUNT=273 GRID=512 bash scripts/gpu_prof.sh
