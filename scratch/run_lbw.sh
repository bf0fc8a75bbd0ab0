#!/bin/bash
# layout backward: column-sum kernel vs ring kernel, layout tests first
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_layout.py tests/test_gpu_baseline_shapes.py -x -q -m gpu -k "layout or canvas or cfg3 or boxes" 2>&1 | tail -8
for sp in 0 4; do
  echo "== colsum splits=$sp"; CSG_BL_ONLY=boxes CSG_LAYOUT_SPLITS=$sp timeout 120 python scratch/bench_layout.py
done
