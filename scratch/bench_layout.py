"""Micro-benchmark of the layout compositor kernels (CUDA events, buffers larger than L2).
    python scratch/bench_layout.py            # cfg2 boxes 64x64 + cfg3 masks 256x256
Prints achieved GB/s = algorithmic bytes (one canvas write / one gradient read) / time."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import synth, _lib          # noqa: E402
from canonicalsg2im_b200.layout import layout_batched  # noqa: E402
from canonicalsg2im_b200.ops import lib, ptr, workspace, _stream  # noqa: E402
from canonicalsg2im_b200.layout import _linspace  # noqa: E402


def objs(seed, n_img, n_min, n_max, D, M):
    vocab = synth.Vocab(0)
    vecs, boxes, masks, off = [], [], [], [0]
    for i in range(n_img):
        g = synth.make_graph(seed * 100 + i, n_min, n_max, vocab, include_dummies=False, mask_size=M)
        n = len(g.boxes)
        vecs.append(synth.det_tensor((n, D), seed * 1000 + i, 1.0))
        boxes.append(g.boxes)
        masks.append(g.masks)
        off.append(off[-1] + n)
    return np.concatenate(vecs), np.concatenate(boxes), np.concatenate(masks), np.array(off, np.int32)


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


def run(tag, N, nmin, nmax, D, H, W, M, use_masks):
    vecs, boxes, masks, off = objs(7, N, nmin, nmax, D, M)
    dev = "cuda"
    v = torch.from_numpy(vecs).to(dev)
    b = torch.from_numpy(boxes).to(dev)
    m = torch.from_numpy(masks).to(dev).float() if use_masks else None
    o = torch.from_numpy(off).to(dev)
    NO = v.shape[0]
    L = lib()
    lin_x, lin_y = _linspace(W, dev), _linspace(H, dev)
    out = torch.empty((N, D, H, W), device=dev)
    Mx = M if use_masks else 0
    mo = int(np.max(np.diff(off)))

    def fwd():
        _lib.check(L.csg_layout_fwd(ptr(v), ptr(b), ptr(m), ptr(o), ptr(lin_x), ptr(lin_y), ptr(out), N, D, H, W, Mx, 0,
                                    mo, _stream()))
    G = torch.randn((N, D, H, W), device=dev)
    dv = torch.empty((NO, D), device=dev)
    ws = workspace(L.csg_layout_bwd_vecs_workspace(N, NO, D, H, W), dev)

    def bwd():
        _lib.check(L.csg_layout_bwd_vecs(ptr(G), ptr(b), ptr(m), ptr(o), ptr(lin_x), ptr(lin_y), ptr(dv), N, NO, D, H, W,
                                         Mx, 0, mo, ptr(ws), ws.numel(), _stream()))
    nbytes = N * D * H * W * 4
    tf, tb = timeit(fwd), timeit(bwd)
    print("%-28s NO=%5d  fwd %7.1f us %7.1f GB/s   bwd %7.1f us %7.1f GB/s" %
          (tag, NO, tf * 1e6, nbytes / tf / 1e9, tb * 1e6, nbytes / tb / 1e9), flush=True)


if __name__ == "__main__":
    _lib.load()
    torch.cuda.set_device(0)
    only = os.environ.get("CSG_BL_ONLY", "")          # "boxes": the boxes_to_layout rows only
    if only != "masks":
        run("cfg2 boxes 128x128x64x64", 128, 3, 31, 128, 64, 64, 16, False)
    if only != "boxes":
        run("cfg3 masks 16x128x256x256", 16, 3, 8, 128, 256, 256, 16, True)
        run("cfg3b masks O=16-24", 16, 16, 24, 128, 256, 256, 16, True)
    if only != "masks":
        run("boxes 32x128x128x128", 32, 3, 31, 128, 128, 128, 16, False)
        run("boxes 8x128x256x256", 8, 3, 31, 128, 256, 256, 16, False)
