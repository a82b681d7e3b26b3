#!/bin/bash
mkdir -p gpurun_out
timeout 200 scripts/kbench/kbench 0 t2 2>&1 | tee gpurun_out/kbench_t2.log
