#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r01g.sh   -- the layout backward after its two-objects-per-iteration rewrite
TAG=r01g
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled \
    -k "regex:layout_bwd_ring_kernel" --launch-skip 0 -c 1 -f -o gpurun_out/prof_${TAG}_k1 \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/ncu_${TAG}_k1.log 2>&1
ls -la gpurun_out | grep ${TAG}
