#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r01d.sh [layout|gemm|all]
# ncu --set full captures (one launch each, warm) of the layout compositor pair on the cfg2 shape and of each
# tcgen05 GEMM variant of the bench step.  Reports land in gpurun_out/ and are summarised into profiles/ here.
TAG=${TAG:-r01d}
WHAT=${1:-all}
mkdir -p gpurun_out
if [ "$WHAT" != "gemm" ]; then
python scratch/bench_layout.py > gpurun_out/bench_layout_${TAG}.log 2>&1
for K in layout_bwd_ring_kernel layout_fwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:${K} --launch-skip 3 -c 1 -f \
      -o gpurun_out/prof_${TAG}_${K} python scratch/bench_layout.py > gpurun_out/ncu_${TAG}_${K}.log 2>&1
done
fi
if [ "$WHAT" != "layout" ]; then
i=0
for K in "gemm_tc_kernel<.int.192, .bool.0, .int.0>" "gemm_tc_kernel<.int.256, .bool.0, .int.0>" "gemm_tc_kernel<.int.256, .bool.1, .int.0>" "gemm_tc_kernel<.int.192, .bool.1, .int.2>" "gemm_tc_kernel<.int.256, .bool.0, .int.1>" "segpool_bf16_kernel<.bool.1>"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${K}" --launch-skip 2 -c 1 -f -o gpurun_out/prof_${TAG}_k${i} \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_k${i}.log 2>&1
done
fi
ls -la gpurun_out | tail -20
