"""Does running the triple-level kernel chain of a layer over graph-aligned CHUNKS of the batch (activations between
consecutive kernels then fit the 126 MB L2) beat one pass over the whole batch?   python scratch/chunked.py"""
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


def time_region(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3    # us


if __name__ == "__main__":
    _lib.load()
    L = lib()
    G, TPG, OPG = 128, 917, 18
    NT, NO, H, Dp = G * TPG, G * OPG, 512, 128
    Wd = 2 * H + Dp
    dev = "cuda"
    hid = rnd((NT, H)); out = rnd((NT, Wd)).abs_()
    w2t = rnd((H, Wd), 0.05); w1t = rnd((384, H), 0.05); w2 = rnd((Wd, H), 0.05); w1 = rnd((H, 384), 0.05)
    b1 = torch.randn(H, device=dev); b2 = torch.randn(Wd, device=dev)
    obj, pred = rnd((NO, 128)), rnd((NT, 128))
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
    valid = torch.ones(NT, dtype=torch.int32, device=dev)
    type32 = torch.ones(NT, dtype=torch.int32, device=dev)
    conf = torch.rand(NT, device=dev)
    dS = torch.randn(NO, H, device=dev); dcnt = torch.randn(NO, device=dev)
    dnewp = rnd((NT, Dp), 0.1)
    pooled32 = torch.empty(NO, H, device=dev); pooled16 = torch.empty(NO, H, device=dev, dtype=BF); cnt = torch.empty(NO, device=dev)
    dobj = torch.empty(NO, 128, device=dev)
    g = torch.empty((NT, Wd), dtype=BF, device=dev); dhid = torch.empty((NT, H), dtype=BF, device=dev)
    dX = torch.empty((NT, 384), dtype=BF, device=dev)
    dconf = torch.empty(NT, device=dev); csg = torch.empty(Wd, device=dev)
    asm_ws = torch.empty(max(L.csg_triple_bwd_assemble_bf16_workspace(NT, H, Dp), 16), dtype=torch.uint8, device=dev)
    dw2 = torch.empty((Wd, H), device=dev); dw1 = torch.empty((H, 384), device=dev)
    ws2 = torch.empty(L.csg_gemm_bf16_workspace(Wd, H, NT, 1), dtype=torch.uint8, device=dev)
    ws1 = torch.empty(L.csg_gemm_bf16_workspace(H, 384, NT, 1), dtype=torch.uint8, device=dev)
    cs_ws = torch.empty(max(L.csg_colsum_bf16_workspace(NT, H), 16), dtype=torch.uint8, device=dev)
    db1 = torch.empty(H, device=dev)
    e = 2    # bytes per bf16

    def st():
        return torch.cuda.current_stream().cuda_stream

    def gemm(mn, gather, M, N, K, A, lda, B, ldb, C, ldc, f32, bias=0, relu=0, rs=0, mask=0, ldm=0, gs=0, go=0, gp=0, ws=0, wsn=0):
        _lib.check(L.csg_gemm_bf16(mn, gather, M, N, K, A, lda, B, ldb, C, ldc, f32, bias, relu, rs, mask, ldm,
                                   ptr(obj) if gather else 0, gp if gather else 0, gs, go, 128 if gather else 0, 128 if gather else 0,
                                   128 if gather else 0, NO if gather else 0, 0, 0, 0, ws, wsn, st()), "gemm")

    def bwd_chain(t0, t1, o0, o1, reuse):
        n = t1 - t0
        # chunk-sized scratch at a fixed address when `reuse` (stays in L2), else the slice of the full buffers
        gq = g.data_ptr() if reuse else g.data_ptr() + t0 * Wd * e
        dh = dhid.data_ptr() if reuse else dhid.data_ptr() + t0 * H * e
        dx = dX.data_ptr() if reuse else dX.data_ptr() + t0 * 384 * e
        _lib.check(L.csg_triple_bwd_assemble_bf16(out.data_ptr() + t0 * Wd * e, ptr(dS), dnewp.data_ptr() + t0 * Dp * e, Dp, ptr(dcnt),
                                                  s.data_ptr() + 4 * t0, o.data_ptr() + 4 * t0, valid.data_ptr() + 4 * t0,
                                                  type32.data_ptr() + 4 * t0, conf.data_ptr() + 4 * t0, n, H, Dp, gq,
                                                  dconf.data_ptr() + 4 * t0, ptr(csg), 0, ptr(asm_ws), asm_ws.numel(), st()), "asm")
        gemm(1, 0, Wd, H, n, gq, Wd, hid.data_ptr() + t0 * H * e, H, ptr(dw2), H, 1, ws=ptr(ws2), wsn=ws2.numel())
        gemm(0, 0, n, H, Wd, gq, Wd, ptr(w2t), Wd, dh, H, 0, mask=hid.data_ptr() + t0 * H * e, ldm=H)
        gemm(1, 2, H, 384, n, dh, H, 0, 0, ptr(dw1), 384, 1, gs=s.data_ptr() + 4 * t0, go=o.data_ptr() + 4 * t0,
             gp=pred.data_ptr() + t0 * 128 * e, ws=ptr(ws1), wsn=ws1.numel())
        _lib.check(L.csg_colsum_bf16(dh, n, H, H, ptr(db1), ptr(cs_ws), cs_ws.numel(), st()), "colsum")
        gemm(0, 0, n, 384, H, dh, H, ptr(w1t), H, dx, 384, 0)
        # gather backward over the objects of the chunk (perm holds global triple ids: base pointer shifted accordingly)
        base = dx - t0 * 384 * e if reuse else dX.data_ptr()
        _lib.check(L.csg_segpool_bf16(base, 384, 0, 256, 128, rps.data_ptr() + 4 * o0, ptr(pms), rpo.data_ptr() + 4 * o0, ptr(pmo), 0, 0,
                                      o1 - o0, dobj.data_ptr() + o0 * 128 * 4, 0, 128, 0, 0, 0, st()), "segpool")

    def fwd_chain(t0, t1, o0, o1, reuse):
        n = t1 - t0
        hq = hid.data_ptr() if reuse else hid.data_ptr() + t0 * H * e
        gemm(0, 1, n, H, 384, 0, 0, ptr(w1), 384, hq, H, 0, bias=ptr(b1), relu=1, gs=s.data_ptr() + 4 * t0, go=o.data_ptr() + 4 * t0,
             gp=pred.data_ptr() + t0 * 128 * e)
        gemm(0, 0, n, Wd, H, hq, H, ptr(w2), H, out.data_ptr() + t0 * Wd * e, Wd, 0, bias=ptr(b2), relu=1, rs=conf.data_ptr() + 4 * t0)
        _lib.check(L.csg_segpool_bf16(ptr(out), Wd, 0, H + Dp, H, rps.data_ptr() + 4 * o0, ptr(pms), rpo.data_ptr() + 4 * o0, ptr(pmo),
                                      ptr(valid), ptr(conf), o1 - o0, pooled32.data_ptr() + o0 * H * 4, pooled16.data_ptr() + o0 * H * e, H,
                                      cnt.data_ptr() + 4 * o0, 1, 0, st()), "pool")

    for name, chain in (("fwd F1,F2,pool", fwd_chain), ("bwd asm,dW2,dhid,dW1,colsum,dX,segpool", bwd_chain)):
        for C in (1, 2, 4, 8, 16):
            gpc = G // C

            def run(reuse):
                for c in range(C):
                    chain(c * gpc * TPG, (c + 1) * gpc * TPG, c * gpc * OPG, (c + 1) * gpc * OPG, reuse)
            for reuse in ((False,) if name.startswith("fwd") and False else (False, True)):
                run(reuse)
                torch.cuda.synchronize()
                cg = torch.cuda.CUDAGraph()
                with torch.cuda.graph(cg):
                    run(reuse)
                t = time_region(cg.replay)
                print("%-42s chunks=%2d reuse_scratch=%d  %8.1f us" % (name, C, int(reuse), t), flush=True)
