#!/bin/bash
for sp in 0 8 16 64; do
  echo "== ring masks splits=$sp"; CSG_BL_ONLY=masks CSG_LAYOUT_SPLITS=$sp timeout 120 python scratch/bench_layout.py
done
