#!/bin/bash
mkdir -p gpurun_out
timeout 300 scripts/kbench/kbench 0 z5 2>&1 | tee gpurun_out/kbench_z5.log
