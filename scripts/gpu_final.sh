#!/bin/bash
# round-1 final single-GPU pass: experiment (software-pipelined mid stages), parity tests, bench, ncu launch list + full capture
set -u
mkdir -p gpurun_out
if [ -x scripts/kbench/kbench_pipe ]; then
  echo "== kbench 1536 strided: baseline vs software-pipelined mid stages"
  timeout 200 scripts/kbench/kbench 1536 s2 2>&1 | grep -v "^== z" | tee gpurun_out/kbench_s2_base.log | tail -22
  timeout 200 scripts/kbench/kbench_pipe 1536 s2 2>&1 | grep -v "^== z" | tee gpurun_out/kbench_s2_pipe.log | tail -22
fi
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/gpu_tests.log
echo "== bench 1024"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_1024_final.log | cut -c1-300
echo "== launch list 1024 (one step after warm-up)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 598 -c 192 --csv \
   --log-file gpurun_out/launches_1024_final.csv python bench.py --grid 1024 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_1024_final.log 2>&1
tail -1 gpurun_out/launches_1024_final.log | cut -c1-200
echo "== ncu --set full z_conv3 @1536: 2nd product of the phi2 group (two operands + accumulate) inside the 2LPT 1024^3 pipeline"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:z_conv3 -s 1 -c 1 \
   -o gpurun_out/prof_z3_1024pipe_final -f python bench.py --grid 1024 --order 2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_z3_1024pipe_final.log 2>&1
tail -2 gpurun_out/prof_z3_1024pipe_final.log | cut -c1-200
