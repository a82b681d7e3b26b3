#!/bin/bash
# One parameterised gpurun call (replaces the one-shot scripts of round 1).  Everything lands in gpurun_out/.
#   TESTS="<pytest -k expr>|all|none"  SMOKE=1  BENCH_GRIDS="1024 512"  BENCH_ARGS="--steps 3 --warmup 3 --no-cpu"
#   LAUNCHES=<grid> (ncu launch list of one step)   NCU_KERNELS="regex ..." NCU_GRID=<grid> (ncu --set full)
#   KBENCH="args" (scripts/kbench/kbench)  KBENCH_ZF="args" (scripts/kbench/kbench_zf)  NG=<gpus> (torchrun for bench when > 1; dist tests)
set -u
TAG=${TAG:-run}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
NG=${NG:-1}
if [ "${TESTS:-none}" != "none" ]; then
  echo "== pytest -m gpu ($TESTS)"
  if [ "$TESTS" = "all" ]; then K=""; else K="-k"; fi
  timeout ${TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q -x --durations=8 $K ${K:+"$TESTS"} 2>&1 | tail -40 | tee gpurun_out/pytest_$TAG.log
fi
if [ "${SMOKE:-0}" = "1" ]; then
  echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5 | tee gpurun_out/smoke_$TAG.log
fi
if [ -n "${KBENCH:-}" ]; then
  echo "== kbench $KBENCH"; timeout 600 scripts/kbench/kbench $KBENCH 2>&1 | tee gpurun_out/kbench_$TAG.log | tail -60
fi
if [ -n "${KBENCH_ZF:-}" ]; then
  echo "== kbench_zf $KBENCH_ZF"; timeout 600 scripts/kbench/kbench_zf $KBENCH_ZF 2>&1 | tee gpurun_out/kbench_zf_$TAG.log | tail -60
fi
for G in ${BENCH_GRIDS:-}; do
  echo "== bench grid $G on $NG GPU(s)"
  if [ "$NG" -gt 1 ]; then
    timeout ${BENCH_TIMEOUT:-900} python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $NG --grid $G ${BENCH_ARGS:---steps 3 --warmup 3 --no-cpu} > gpurun_out/bench_${TAG}_n${NG}_$G.log 2>&1
  else
    timeout ${BENCH_TIMEOUT:-900} python bench.py --grid $G ${BENCH_ARGS:---steps 3 --warmup 3 --no-cpu} > gpurun_out/bench_${TAG}_n${NG}_$G.log 2>&1
  fi
  tail -1 gpurun_out/bench_${TAG}_n${NG}_$G.log | cut -c1-1500
done
if [ -n "${LAUNCHES:-}" ]; then
  echo "== ncu launch list grid $LAUNCHES"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s ${LSKIP:-0} -c ${LCOUNT:-20000} --csv \
     --log-file gpurun_out/launches_${TAG}_$LAUNCHES.csv python bench.py --grid $LAUNCHES --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_${TAG}_$LAUNCHES.log 2>&1
  tail -1 gpurun_out/launches_${TAG}_$LAUNCHES.log | cut -c1-200
fi
for K in ${NCU_KERNELS:-}; do
  echo "== ncu --set full $K"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s ${FSKIP:-40} -c ${FCOUNT:-3} \
     -o gpurun_out/prof_${TAG}_${K} -f python bench.py --grid ${NCU_GRID:-1024} --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_${TAG}_${K}.log 2>&1
  tail -1 gpurun_out/prof_${TAG}_${K}.log | cut -c1-200
done
ls -la gpurun_out | tail -5
