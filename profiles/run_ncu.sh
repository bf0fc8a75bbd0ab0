#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu.sh <tag> [extra bench args]
# 1) launch list with per-launch device time for a short bench run; 2) full-section captures of the
# dominant kernels (one launch each).  Outputs go to gpurun_out/ and are summarised into profiles/.
TAG=${1:-r01}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in layout_fwd_kernel layout_bwd_vecs_kernel segpool_kernel gemm_f32_kernel gemm_tc triple_bwd_assemble canon_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 6 -c 1 -f -o gpurun_out/prof_${TAG}_${K} \
      python bench.py --steps 1 --warmup 1 --no-cpu-baseline "$@" > gpurun_out/ncu_${TAG}_${K}.log 2>&1
done
ls -la gpurun_out | tail -20
