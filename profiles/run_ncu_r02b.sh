#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r02b.sh [tag]
# ncu --set full captures (one launch each) of the HBM-bound kernels around net1 and of the dominant GEMM
TAG=${1:-r02b}
mkdir -p gpurun_out
i=0
for K in "triple_bwd_assemble_bf16_kernel" "segpool_bf16_kernel<.bool.1>" "colsum_bf16_partial_kernel" "gemm_tc_kernel<.int.256, .bool.0, .int.0, .int.1, .int.1>"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${K}" --launch-skip 2 -c 1 -f -o gpurun_out/prof_${TAG}_k${i} \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_k${i}.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
