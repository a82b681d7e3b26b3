#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== new GPU tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "every_long_plan or convolver or pipeline or grouping or generations" 2>&1 | tail -4 | tee gpurun_out/gpu_tests_last.log
echo "== bench"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_last.log | cut -c1-300
