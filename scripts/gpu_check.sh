#!/bin/bash
# One gpurun call: GPU parity tests, smoke, short benches.  Outputs land in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 2>&1 | tail -40 | tee gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke.log
for G in ${GRIDS:-256 512}; do
  echo "== bench grid $G"; timeout 900 python bench.py --grid $G --steps 3 --warmup 3 --cpu-grid 64 2>&1 | tail -3 | tee gpurun_out/bench_$G.log
done
