#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r01f.sh      (final state of round 1)
# 1) launch list (per-launch device time, cold-cache, serialised) of one profiled bench step;
# 2) ncu --set full captures (one launch each) of the dominant GEMM (F2) and of the layout compositor pair.
TAG=r01f
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
i=0
for K in "gemm_tc_kernel<.int.256, .bool.0, .int.0, .int.1, .int.1>" "layout_fwd_kernel" "layout_bwd_ring_kernel" "gemm_tc_kernel<.int.128, .bool.1, .int.2, .int.1, .int.2>"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${K}" --launch-skip 0 -c 1 -f -o gpurun_out/prof_${TAG}_k${i} \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_k${i}.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
