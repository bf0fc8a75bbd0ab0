"""add_location_triplets: device kernel (whole batch) vs the oracle port of the reference's Python loops (per graph).
    python scratch/bench_location.py"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import synth, _lib
from canonicalsg2im_b200 import canonicalize as C
from oracle import canon as ocanon

_lib.load()
vocab = synth.Vocab(42)
for n_min, n_max, B in ((3, 30, 128), (32, 64, 10)):
    graphs = [synth.make_graph(100 + i, n_min, n_max, vocab, include_dummies=True) for i in range(B)]
    cens = [np.concatenate([g.centers, np.zeros((len(g.boxes) - len(g.centers), 2), np.float32)]).astype(np.float32) for g in graphs]
    boxes = torch.from_numpy(np.concatenate([g.boxes for g in graphs]).astype(np.float32)).cuda()
    cen = torch.from_numpy(np.concatenate(cens)).cuda()
    objs = torch.from_numpy(np.concatenate([g.objs for g in graphs]).astype(np.int64)).cuda()
    off = torch.from_numpy(np.concatenate([[0], np.cumsum([len(g.boxes) for g in graphs])]).astype(np.int32)).cuda()
    mo = max(len(g.boxes) for g in graphs)
    for _ in range(3):
        C.add_location_triplets_batched(boxes, cen, objs, off, vocab.image_obj_id, vocab.pred_ids, mo)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        trip, _ = C.add_location_triplets_batched(boxes, cen, objs, off, vocab.image_obj_id, vocab.pred_ids, mo)
    torch.cuda.synchronize()
    gpu = (time.perf_counter() - t0) / 20
    t0 = time.perf_counter()
    for g, c in zip(graphs[:8], cens[:8]):
        ocanon.add_location_triplets(g.boxes, c, g.objs[:, 0], vocab.image_obj_id, vocab.pred_ids)
    cpu = (time.perf_counter() - t0) / 8
    print("objects %d-%d, %d graphs: device %.3f ms per batch (%d triplets, incl. the size read-back) | oracle port %.2f ms per graph"
          % (n_min, n_max, B, gpu * 1e3, trip.shape[0], cpu * 1e3), flush=True)
