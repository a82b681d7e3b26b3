#!/bin/bash
mkdir -p gpurun_out
timeout 200 scripts/kbench/kbench 1536 seg 2>&1 | tee gpurun_out/kbench_seg2.log
