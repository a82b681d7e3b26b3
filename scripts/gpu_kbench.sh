#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== kbench"; timeout 600 scripts/kbench/kbench ${KARGS:-0 zs} 2>&1 | tee gpurun_out/kbench_zs.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench 1024"; timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e 2>&1 | tail -1 | tee gpurun_out/bench_1024_tw2.log | cut -c1-300
