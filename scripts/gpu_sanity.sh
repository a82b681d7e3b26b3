#!/bin/bash
# end-of-round sanity on one GPU: smoke(), the GPU parity suite, the default bench line and the reference arm
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.log
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gpu_tests.log
echo "== bench (default flags)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.log | cut -c1-300
echo "== bench --impl reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-400
