#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu.sh <tag> <kernel-regex-list> [extra bench args]
# 1) launch list with per-launch device time for one profiled bench step; 2) full-section captures of the
# named kernels (one launch each).  Outputs go to gpurun_out/ and are summarised into profiles/ by summarize.py.
TAG=${1:-r01}; shift
KERNELS=${1:-"gemm_tc_kernel layout_fwd_kernel layout_bwd_vecs_kernel segpool_bf16_kernel triple_bwd_assemble_bf16_kernel canon_kernel"}; shift
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --profile "$@" > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
for K in $KERNELS; do
  ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:${K} -c 1 -f \
      -o gpurun_out/prof_${TAG}_${K} \
      python bench.py --steps 1 --warmup 3 --profile "$@" > gpurun_out/ncu_${TAG}_${K}.log 2>&1
done
ls -la gpurun_out | tail -20
