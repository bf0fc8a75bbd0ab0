#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r02a.sh
# launch list (per-launch device time, cold-cache, serialised) of one profiled bench step of the round-2 state
TAG=${1:-r02a}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 1 --warmup 3 --profile > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ls -la gpurun_out | grep ${TAG}
