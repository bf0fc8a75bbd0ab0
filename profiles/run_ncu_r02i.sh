#!/bin/bash
# Usage (under gpurun): bash profiles/run_ncu_r02i.sh
# Final state of round 2 (third session): ncu launch list of one profiled cfg2 step (the column-sum layout backward in place)
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv \
    --log-file gpurun_out/launches_r02i.csv python bench.py --steps 1 --warmup 3 --profile > gpurun_out/bench_under_ncu_r02i.log 2>&1
tail -2 gpurun_out/bench_under_ncu_r02i.log | cut -c1-300
wc -l gpurun_out/launches_r02i.csv
