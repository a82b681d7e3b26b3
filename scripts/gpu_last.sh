#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.log
echo "== GPU tests (long plans, convolver, pipeline, grouping, generations)"; timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "every_long_plan or convolver or pipeline or grouping or generations" 2>&1 | tail -4 | tee gpurun_out/gpu_tests_last.log
