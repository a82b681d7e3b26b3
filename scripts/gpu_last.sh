#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench (default flags)"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.log | cut -c1-300
