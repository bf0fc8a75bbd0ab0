#!/bin/bash
# A/B of the gradient all-reduce launch grouping (CSG_GRAD_GROUPS) at N = 8 (and N = 1) on one box
mkdir -p gpurun_out
N=${N:-8}
python bench.py --gpus 1 --steps 30 --warmup 5 --configs none > gpurun_out/groups_n1.json 2> gpurun_out/groups_n1.err
python -c "
import json; d = json.loads(open('gpurun_out/groups_n1.json').read().strip().splitlines()[-1]); print('n1', d['ms_per_step'], d['value'])"
for mode in layers end two; do
  CSG_GRAD_GROUPS=$mode timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29500 bench.py --gpus $N --steps 30 --warmup 5 --configs none > gpurun_out/groups_${mode}_n$N.json 2> gpurun_out/groups_${mode}_n$N.err
  echo "$mode rc=$?"
  python -c "
import json; d = json.loads(open('gpurun_out/groups_${mode}_n$N.json').read().strip().splitlines()[-1]); print('$mode', d['ms_per_step'], d['value'], 'e2e', d['e2e']['ms_per_step'])"
done
