#!/bin/bash
# second pass of r01e: the F2 GEMM (first launch of its template in the step) and the layout compositor pair
TAG=r01e
mkdir -p gpurun_out
i=0
for K in "gemm_tc_kernel<.int.256, .bool.0, .int.0, .int.1>" "layout_fwd_kernel" "layout_bwd_ring_kernel"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${K}" --launch-skip 0 -c 1 -f -o gpurun_out/prof_${TAG}_b${i} \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_b${i}.log 2>&1
done
ls -la gpurun_out | grep "_b[0-9]"
