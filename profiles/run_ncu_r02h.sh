#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r02h.sh
# ncu --set full of the column-sum layout backward (boxes_to_layout, cfg2 canvas 128 x 128 x 64 x 64) from the layout micro-benchmark
mkdir -p gpurun_out
CSG_BL_ONLY=boxes ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k "regex:layout_bwd_colsum_kernel" --launch-skip 5 -c 1 -f -o gpurun_out/prof_r02h_k1 \
    python scratch/bench_layout.py > gpurun_out/ncu_r02h_k1.log 2>&1
tail -3 gpurun_out/ncu_r02h_k1.log
ls -la gpurun_out | grep r02h
