"""Do the memory-bound kernels of a layer's backward overlap with a persistent tcgen05 GEMM launched on another stream?
    python scratch/corun.py
Prints, for pairs (GEMM, other): each alone, back to back on one stream, and on two streams (fork / join by events)."""
import ctypes
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import _lib, ops  # noqa: E402
from canonicalsg2im_b200.ops import lib, ptr    # noqa: E402
from canonicalsg2im_b200 import graph_tc   # noqa: E402

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
    NT, NO, H, Dp = 117321, 2349, 512, 128
    Wd = 2 * H + Dp
    dev = "cuda"
    hid = rnd((NT, H)); g = rnd((NT, Wd), 0.1); dhid = rnd((NT, H), 0.1); out = rnd((NT, Wd)).abs_()
    w2t = rnd((H, Wd), 0.05); w1t = rnd((384, H), 0.05); w2 = rnd((Wd, H), 0.05)
    obj, pred = rnd((NO, 128)), rnd((NT, 128))
    # VG-like structure: ~18 objects and ~900 triples per graph
    gid = torch.arange(NT, device=dev) // 917
    s = (gid * 18 + torch.randint(0, 18, (NT,), device=dev)).clamp_(max=NO - 1).int()
    o = (gid * 18 + torch.randint(0, 18, (NT,), device=dev)).clamp_(max=NO - 1).int()
    ga = ops.Gather(obj, pred, s, o)

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
    gout = torch.empty((NT, Wd), dtype=BF, device=dev)
    dconf = torch.empty(NT, device=dev); csg = torch.empty(Wd, device=dev)
    asm_ws = torch.empty(max(L.csg_triple_bwd_assemble_bf16_workspace(NT, H, Dp), 16), dtype=torch.uint8, device=dev)
    dw2 = torch.empty((Wd, H), device=dev); dw1 = torch.empty((H, 384), device=dev)
    outh = torch.empty((NT, H), dtype=BF, device=dev); out2 = torch.empty((NT, Wd), dtype=BF, device=dev)
    outx = torch.empty((NT, 384), dtype=BF, device=dev)
    ws2 = torch.empty(L.csg_gemm_bf16_workspace(Wd, H, NT, 1), dtype=torch.uint8, device=dev)
    ws1 = torch.empty(L.csg_gemm_bf16_workspace(H, 384, NT, 1), dtype=torch.uint8, device=dev)
    cs_ws = torch.empty(max(L.csg_colsum_bf16_workspace(NT, H), 16), dtype=torch.uint8, device=dev)
    db1 = torch.empty(H, device=dev)

    def st():
        return torch.cuda.current_stream().cuda_stream

    def k_dw2():
        _lib.check(L.csg_gemm_bf16(1, 0, Wd, H, NT, ptr(g), Wd, ptr(hid), H, ptr(dw2), H, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                   0, 0, 0, 0, ptr(ws2), ws2.numel(), st()), "dw2")

    def k_dw1():
        _lib.check(L.csg_gemm_bf16(1, 2, H, 384, NT, ptr(dhid), H, 0, 0, ptr(dw1), 384, 1, 0, 0, 0, 0, 0, ptr(obj), ptr(pred),
                                   ptr(s), ptr(o), 128, 128, 128, NO, 0, 0, 0, ptr(ws1), ws1.numel(), st()), "dw1")

    def k_dhid():
        ops.gemm_bf16(NT, H, Wd, g, w2t, mask_aux=hid, out=outh)

    def k_dx():
        ops.gemm_bf16(NT, 384, H, dhid, w1t, out=outx)

    def k_f2():
        ops.gemm_bf16(NT, Wd, H, hid, w2, relu=True, rowscale=conf, out=out2)

    def k_asm():
        _lib.check(L.csg_triple_bwd_assemble_bf16(ptr(out), ptr(dS), ptr(dnewp), Dp, ptr(dcnt), ptr(s), ptr(o), ptr(valid),
                                                  ptr(type32), ptr(conf), NT, H, Dp, ptr(gout), ptr(dconf), ptr(csg), 0,
                                                  ptr(asm_ws), asm_ws.numel(), st()), "asm")

    def k_pool():
        _lib.check(L.csg_segpool_bf16(ptr(out), Wd, 0, H + Dp, H, ptr(rps), ptr(pms), ptr(rpo), ptr(pmo), ptr(valid), ptr(conf),
                                      NO, ptr(pooled32), ptr(pooled16), H, ptr(cnt), 1, 0, st()), "pool")

    def k_colsum():
        _lib.check(L.csg_colsum_bf16(ptr(dhid), NT, H, H, ptr(db1), ptr(cs_ws), cs_ws.numel(), st()), "colsum")

    side = torch.cuda.Stream()
    ev_fork, ev_join = torch.cuda.Event(), torch.cuda.Event()

    def corun(a, b, first_side=False):
        def fn():
            main = torch.cuda.current_stream()
            ev_fork.record(main)
            side.wait_event(ev_fork)
            if first_side:
                with torch.cuda.stream(side):
                    b()
                a()
            else:
                a()
                with torch.cuda.stream(side):
                    b()
            ev_join.record(side)
            main.wait_event(ev_join)
        return fn

    def k_torch():
        gout.mul_(1.0001)

    singles = {"torchmul": k_torch, "dW2": k_dw2, "dW1": k_dw1, "dhid": k_dhid, "dX": k_dx, "F2": k_f2, "assemble": k_asm, "pool": k_pool, "colsum": k_colsum}
    t1 = {}
    for n, f in singles.items():
        t1[n] = time_region(f)
        print("alone   %-10s %7.1f us" % (n, t1[n]), flush=True)
    pairs = [("dW2", "torchmul"), ("F2", "torchmul"), ("dW2", "assemble"), ("dW2", "pool"), ("dW2", "colsum"), ("dW1", "assemble"), ("dW1", "pool"),
             ("dhid", "assemble"), ("F2", "pool"), ("dX", "colsum"), ("dW2", "dhid"), ("dW2", "dX")]
    for a, b in pairs:
        fa, fb = singles[a], singles[b]
        ser = time_region(lambda: (fa(), fb()))
        co = time_region(corun(fa, fb))
        co2 = time_region(corun(fa, fb, first_side=True))
        print("pair %-6s + %-9s  sum-alone %7.1f  serial %7.1f  two-streams(a first) %7.1f  (b first) %7.1f" %
              (a, b, t1[a] + t1[b], ser, co, co2), flush=True)
