"""Cost of the segment-sum formulation of net1's first-layer backward (dW1 / dX / gather backward / db1 from per-object
sums of dhid) against the current gathered-B dW1 + full dX + segpool + colsum.   python scratch/bwd_alt.py"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import _lib, ops  # noqa: E402
from canonicalsg2im_b200.ops import lib, ptr    # noqa: E402

BF = torch.bfloat16


def rnd(shape, scale=1.0):
    return (torch.randn(shape, device="cuda") * scale).to(BF)


def time_graph(fn, iters=20):
    fn()
    torch.cuda.synchronize()
    cg = torch.cuda.CUDAGraph()
    with torch.cuda.graph(cg):
        fn()
    for _ in range(3):
        cg.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        cg.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


if __name__ == "__main__":
    _lib.load()
    L = lib()
    G, TPG, OPG = 128, 917, 18
    NT, NO, H, D = G * TPG, G * OPG, 512, 128
    dev = "cuda"
    # the producer of dhid in the step is a GEMM over 270 MB of g: emulate its L2 footprint with a big copy before
    big = torch.empty(64 * 1024 * 1024, device=dev); big2 = torch.empty_like(big)
    dhid = rnd((NT, H), 0.1)
    w1t = rnd((384, H), 0.05); wso = rnd((D, 2 * H), 0.05)
    obj, pred = rnd((NO, D)), rnd((NT, D))
    gid = torch.arange(NT, device=dev) // TPG
    s = (gid * OPG + torch.randint(0, OPG, (NT,), device=dev)).int()
    o = (gid * OPG + torch.randint(0, OPG, (NT,), device=dev)).int()

    def csr(idx):
        perm = torch.sort(idx.long(), stable=True)[1].int()
        cnt = torch.bincount(idx.long(), minlength=NO)
        rowptr = torch.zeros(NO + 1, dtype=torch.int32, device=dev)
        rowptr[1:] = torch.cumsum(cnt, 0).int()
        return rowptr, perm
    rps, pms = csr(s); rpo, pmo = csr(o)
    zero_rp = torch.zeros(NO + 1, dtype=torch.int32, device=dev)
    dHso = torch.empty((NO, 2 * H), dtype=BF, device=dev)
    dX = torch.empty((NT, 384), dtype=BF, device=dev); dXp = torch.empty((NT, D), dtype=BF, device=dev)
    dobj = torch.empty((NO, D), device=dev)
    dw1 = torch.empty((H, 384), device=dev); dw1p = torch.empty((H, D), device=dev); dw1so = torch.empty((2 * H, D), device=dev)
    db1 = torch.empty(H, device=dev)

    def ws(M, N, K):
        return torch.empty(max(L.csg_gemm_bf16_workspace(M, N, K, 1), 16), dtype=torch.uint8, device=dev)
    ws1, wsp, wsso = ws(H, 384, NT), ws(H, D, NT), ws(2 * H, D, NO)
    cs_ws = torch.empty(max(L.csg_colsum_bf16_workspace(NT, H), 16), dtype=torch.uint8, device=dev)

    def st():
        return torch.cuda.current_stream().cuda_stream

    def gemm(mn, gather, M, N, K, A, lda, B, ldb, C, ldc, f32, w=None):
        _lib.check(L.csg_gemm_bf16(mn, gather, M, N, K, A, lda, B, ldb, C, ldc, f32, 0, 0, 0, 0, 0,
                                   ptr(obj) if gather else 0, ptr(pred) if gather else 0, ptr(s) if gather else 0, ptr(o) if gather else 0,
                                   D if gather else 0, D if gather else 0, D if gather else 0, NO if gather else 0, 0, 0, 0,
                                   ptr(w), w.numel() if w is not None else 0, st()), "gemm")

    def segsum(rp_s, rp_o, X, ldx, cs, co, W, out32, out16, ldo):
        _lib.check(L.csg_segpool_bf16(X, ldx, cs, co, W, ptr(rp_s), ptr(pms), ptr(rp_o), ptr(pmo), 0, 0, NO, out32, out16, ldo, 0, 0, 0, st()), "seg")

    def flush():
        big2.copy_(big)

    parts = {
        "flush (copy 256 MB -> subtract)": lambda: flush(),
        "old dW1 gathered": lambda: (flush(), gemm(1, 2, H, 384, NT, ptr(dhid), H, 0, 0, ptr(dw1), 384, 1, ws1)),
        "old colsum(dhid)": lambda: (flush(), _lib.check(L.csg_colsum_bf16(ptr(dhid), NT, H, H, ptr(db1), ptr(cs_ws), cs_ws.numel(), st()), "cs")),
        "old dX N=384": lambda: (flush(), gemm(0, 0, NT, 384, H, ptr(dhid), H, ptr(w1t), H, ptr(dX), 384, 0)),
        "old segpool(dX)": lambda: (flush(), segsum(rps, rpo, ptr(dX), 384, 0, 256, D, ptr(dobj), 0, D)),
        "new segsum s (dhid)": lambda: (flush(), segsum(rps, zero_rp, ptr(dhid), H, 0, 0, H, 0, ptr(dHso), 2 * H)),
        "new segsum s+o (dhid)": lambda: (flush(), segsum(rps, zero_rp, ptr(dhid), H, 0, 0, H, 0, ptr(dHso), 2 * H),
                                          segsum(zero_rp, rpo, ptr(dhid), H, 0, 0, H, 0, dHso.data_ptr() + 2 * H, 2 * H)),
        "new colsum(dHs)": lambda: (flush(), _lib.check(L.csg_colsum_bf16(ptr(dHso), NO, H, 2 * H, ptr(db1), ptr(cs_ws), cs_ws.numel(), st()), "cs")),
        "new dXp N=128": lambda: (flush(), gemm(0, 0, NT, D, H, ptr(dhid), H, w1t.data_ptr() + 128 * H * 2, H, ptr(dXp), D, 0)),
        "new dW1p 512x128xT": lambda: (flush(), gemm(1, 0, H, D, NT, ptr(dhid), H, ptr(pred), D, ptr(dw1p), D, 1, wsp)),
        "new dW1so 1024x128xNO": lambda: (flush(), gemm(1, 0, 2 * H, D, NO, ptr(dHso), 2 * H, ptr(obj), D, ptr(dw1so), D, 1, wsso)),
        "new dobj NOx128x1024": lambda: (flush(), gemm(0, 0, NO, D, 2 * H, ptr(dHso), 2 * H, ptr(wso), 2 * H, ptr(dobj), D, 1)),
    }
    t_flush = None
    for name, fn in parts.items():
        t = time_graph(fn)
        if t_flush is None:
            t_flush = t
            print("%-36s %8.1f us" % (name, t))
        else:
            print("%-36s %8.1f us" % (name, t - t_flush), flush=True)

    def old_chain():
        gemm(1, 2, H, 384, NT, ptr(dhid), H, 0, 0, ptr(dw1), 384, 1, ws1)
        _lib.check(L.csg_colsum_bf16(ptr(dhid), NT, H, H, ptr(db1), ptr(cs_ws), cs_ws.numel(), st()), "cs")
        gemm(0, 0, NT, 384, H, ptr(dhid), H, ptr(w1t), H, ptr(dX), 384, 0)
        segsum(rps, rpo, ptr(dX), 384, 0, 256, D, ptr(dobj), 0, D)

    def new_chain():
        segsum(rps, zero_rp, ptr(dhid), H, 0, 0, H, 0, ptr(dHso), 2 * H)
        segsum(zero_rp, rpo, ptr(dhid), H, 0, 0, H, 0, dHso.data_ptr() + 2 * H, 2 * H)
        _lib.check(L.csg_colsum_bf16(ptr(dHso), NO, H, 2 * H, ptr(db1), ptr(cs_ws), cs_ws.numel(), st()), "cs")
        gemm(1, 0, H, D, NT, ptr(dhid), H, ptr(pred), D, ptr(dw1p), D, 1, wsp)
        gemm(1, 0, 2 * H, D, NO, ptr(dHso), 2 * H, ptr(obj), D, ptr(dw1so), D, 1, wsso)
        gemm(0, 0, NT, D, H, ptr(dhid), H, w1t.data_ptr() + 128 * H * 2, H, ptr(dXp), D, 0)
        gemm(0, 0, NO, D, 2 * H, ptr(dHso), 2 * H, ptr(wso), 2 * H, ptr(dobj), D, 1)
    print("old chain (dW1, colsum, dX, segpool)      %8.1f us" % (time_graph(lambda: (flush(), old_chain())) - t_flush))
    print("new chain (segsum x2, colsum, 4 GEMMs)    %8.1f us" % (time_graph(lambda: (flush(), new_chain())) - t_flush))
