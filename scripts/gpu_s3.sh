#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== kbench s3"; timeout 600 scripts/kbench/kbench 0 s3 2>&1 | tee gpurun_out/kbench_s3.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench 1024"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1024_v3.log | cut -c1-400
