#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r02g.sh [tag]
# Final state of round 2: launch list of one profiled step + ncu --set full captures of the kernels that carry the step:
# F2 (K-major GEMM with the staged bf16 epilogue), the CTA-pair dhid GEMM, the split-K dW2 GEMM (third launch of that
# instantiation in the backward: the first two are net2's small weight gradients), the pipelined assemble kernel and the
# per-object segment sums of dhidden.
TAG=${1:-r02g}
mkdir -p gpurun_out
if [ -z "$ONLY" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 1 --warmup 3 --profile > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
fi
i=0
# demangled names carry the parameter types: gemm_tc_kernel<(int)256, (bool)0, (int)0, (int)1, (int)1, (int)8, (bool)0>
G="gemm_tc_kernel<.int.256, .bool."
for SPEC in "${G}0, .int.0, .int.1, .int.1, .int.8, .bool.0>:0" "${G}0, .int.0, .int.2, .int.1, .int.8, .bool.0>:0" \
            "${G}1, .int.0, .int.1, .int.1, .int.8, .bool.0>:2" "triple_bwd_assemble_pipe_kernel:0" \
            "segpool_bf16_kernel<.bool.0>:0" "segpool_bf16_kernel<.bool.1>:0"; do
  i=$((i+1))
  K="${SPEC%%:*}"; SKIP="${SPEC##*:}"
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $i "; then continue; fi
  ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
      -k "regex:${K}" --launch-skip ${SKIP} -c 1 -f -o gpurun_out/prof_${TAG}_k${i} \
      python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_k${i}.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
