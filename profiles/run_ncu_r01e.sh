#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r01e.sh
# 1) launch list (per-launch device time, cold-cache, serialised) of one profiled bench step;
# 2) ncu --set full captures (one warm launch each) of the dominant kernels of the step.
# Reports land in gpurun_out/ and are summarised into profiles/r01e_summary.md by profiles/summarize.py r01e.
TAG=r01e
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
i=0
for K in "gemm_tc_kernel<.int.256, .bool.0, .int.0, .int.1>" "gemm_tc_kernel<.int.256, .bool.0, .int.0, .int.2>" "gemm_tc_kernel<.int.256, .bool.1, .int.0, .int.1>" "gemm_tc_kernel<.int.256, .bool.0, .int.1, .int.1>" "triple_bwd_assemble_bf16_kernel" "segpool_bf16_kernel<.bool.1>" "layout_fwd_kernel" "layout_bwd_ring_kernel"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${K}" --launch-skip 1 -c 1 -f -o gpurun_out/prof_${TAG}_k${i} \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_k${i}.log 2>&1
done
ls -la gpurun_out | tail -14
