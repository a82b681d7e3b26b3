#!/bin/bash
# ncu passes for one bench config (1 GPU).  GRID defaults to 512.
set -u
G=${GRID:-512}
mkdir -p gpurun_out
echo "== launch list (gpu__time_duration, one step after warm-up)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-1100} -c ${COUNT:-300} --csv \
   --log-file gpurun_out/launches_$G.csv python bench.py --grid $G --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_$G.log 2>&1
tail -2 gpurun_out/launches_$G.log | cut -c1-300
for K in ${KERNELS:-strided_pass_kernel z_conv_kernel}; do
  echo "== ncu --set full $K"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${FSKIP:-40} -c ${FCOUNT:-6} \
     -o gpurun_out/prof_${K}_$G -f python bench.py --grid $G --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_${K}_$G.log 2>&1
  tail -2 gpurun_out/prof_${K}_$G.log | cut -c1-200
done
ls -la gpurun_out
