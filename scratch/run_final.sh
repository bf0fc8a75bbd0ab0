#!/bin/bash
mkdir -p gpurun_out
TAG=${TAG:-r2l}
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${TAG}_n1.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["frac"], "hbm fwd", d["roofline_hbm"]["fwd"], "bwd", d["roofline_hbm"]["bwd"])
for k, v in d["configs"].items():
    print(k, {kk: vv for kk, vv in v.items() if kk in ("ms_per_step", "ms_fwd_bwd", "graphs_per_s", "error")})
print("cfg1", d["configs"]["cfg1"].get("bf16", {}).get("ms_per_step"), d["configs"]["cfg1"].get("roofline_hbm", {}).get("bwd"))
PY
CSG_BL_ONLY=boxes timeout 120 python scratch/bench_layout.py
