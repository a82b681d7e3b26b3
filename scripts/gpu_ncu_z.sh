#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:z_conv3 -c 2 -o gpurun_out/prof_z3_768 -f scripts/kbench/kbench 768 z3 > gpurun_out/prof_z3_768.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:z_conv3 -c 2 -o gpurun_out/prof_z3_1536 -f scripts/kbench/kbench 1536 z3 > gpurun_out/prof_z3_1536.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strided2 -c 3 -o gpurun_out/prof_s2_1536 -f scripts/kbench/kbench 1536 s > gpurun_out/prof_s2_1536.log 2>&1
ls -la gpurun_out/*.ncu-rep
