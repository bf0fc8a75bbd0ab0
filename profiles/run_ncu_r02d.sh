#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r02d.sh [tag]
# ncu --set full captures of the canonicalization kernels and of the reworked assemble / pooling kernels
TAG=${1:-r02d}
mkdir -p gpurun_out
i=0
for K in "canon_kernel<.bool.0>" "canon_kernel<.bool.1>" "triple_bwd_assemble_bf16_kernel" "segpool_bf16_kernel<.bool.1>"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${K}" --launch-skip 0 -c 1 -f -o gpurun_out/prof_${TAG}_k${i} \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_k${i}.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
