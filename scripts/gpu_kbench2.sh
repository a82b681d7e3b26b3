#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== seg 1536"; timeout 300 scripts/kbench/kbench 1536 seg 2>&1 | tee gpurun_out/kbench_seg.log
echo "== z4"; timeout 600 scripts/kbench/kbench 0 z4 2>&1 | tee gpurun_out/kbench_z4.log
echo "== pytest gpu (new tests)"; timeout 900 python -m pytest tests -m gpu -q -x -k "naive or ppt or complex or stream" 2>&1 | tail -5 | tee gpurun_out/pytest_gpu_new.log
