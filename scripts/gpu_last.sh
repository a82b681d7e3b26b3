#!/bin/bash
mkdir -p gpurun_out
timeout 50 python -m pytest tests/test_gpu_parity.py -x -q -k "grouping" 2>&1 | tail -3 | tee gpurun_out/gpu_tests_grouping.log
