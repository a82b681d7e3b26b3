#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_cpp_driver.py tests/test_cpp_mirror.py -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/gpu_tests_cpp.log
