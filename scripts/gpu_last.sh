#!/bin/bash
mkdir -p gpurun_out
timeout 25 scripts/kbench/kbench 0 zacc 2>&1 | tee gpurun_out/kbench_zacc.log
