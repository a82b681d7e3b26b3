#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== kbench s3"; timeout 600 scripts/kbench/kbench 0 s3 2>&1 | tee gpurun_out/kbench_s3b.log
