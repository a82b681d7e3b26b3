#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== kbench z4 3072 + acc"; timeout 300 scripts/kbench/kbench 3072 z4 2>&1 | tee gpurun_out/kbench_z4_3072.log; timeout 200 scripts/kbench/kbench 1536 z4 2>&1 | tail -3 | tee -a gpurun_out/kbench_z4_3072.log
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "== bench 1024"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1024_v5.log | cut -c1-300
echo "== ncu --set full z_conv3 @1536 (2nd product of the phi2 group: two operands + accumulate) inside the 2LPT 1024^3 pipeline"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:z_conv3 -s 4 -c 1 \
   -o gpurun_out/prof_z3_1024pipe_v5 -f python bench.py --grid 1024 --order 2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_z3_1024pipe_v5.log 2>&1
tail -2 gpurun_out/prof_z3_1024pipe_v5.log | cut -c1-200
