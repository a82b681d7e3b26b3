#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_parity.py -x -q -k two_fluid 2>&1 | tail -3 | tee gpurun_out/gpu_tests_two_fluid.log
