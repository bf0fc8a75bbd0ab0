#!/bin/bash
# A/B of the gradient all-reduce launch grouping (CSG_GRAD_GROUPS) at N = 2 on one box
mkdir -p gpurun_out
N=${N:-2}
for mode in layers end two; do
  extra="--configs none"
  [ "$mode" = "two" ] && extra="--configs cfg5,nccl_check"
  CSG_GRAD_GROUPS=$mode timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29500 bench.py --gpus $N --steps 20 --warmup 5 $extra > gpurun_out/groups_${mode}_n$N.json 2> gpurun_out/groups_${mode}_n$N.err
  echo "$mode rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/groups_${mode}_n$N.json").read().strip().splitlines()[-1])
    print("$mode", d["ms_per_step"], d["value"], "e2e", d["e2e"]["ms_per_step"])
    c5 = d.get("configs", {}).get("cfg5")
    if c5: print(" cfg5", {k: c5.get(k) for k in ("ms_per_step", "graphs_per_s", "allreduce_tail_ms", "nccl_check", "error")})
except Exception as ex:
    print("$mode parse failed", ex)
PY
done
