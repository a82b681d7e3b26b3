#!/bin/bash
# round 1, session 7: tests + bench (streamed hand-off) + 1024^3 launch list + ncu --set full of the fused z pass at 1536
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | tee gpurun_out/gpu.txt
free -g | head -2 | tee -a gpurun_out/gpu.txt; nproc | tee -a gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/gpu_tests.log
echo "== bench 1024"; timeout 900 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_1024.log
echo "== launch list 1024 (one step after warm-up)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 850 -c 300 --csv \
   --log-file gpurun_out/launches_1024.csv python bench.py --grid 1024 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_1024.log 2>&1
tail -1 gpurun_out/launches_1024.log | cut -c1-200
echo "== ncu --set full z_conv3 @1536 inside the 2LPT 1024^3 pipeline"
timeout 500 ncu --set full --clock-control none --import-source on -k regex:z_conv3 -s 6 -c 1 \
   -o gpurun_out/prof_z3_1024pipe -f python bench.py --grid 1024 --order 2 --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_z3_1024pipe.log 2>&1
tail -2 gpurun_out/prof_z3_1024pipe.log | cut -c1-200
ls -la gpurun_out | head -30
