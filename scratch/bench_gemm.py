"""Micro-benchmark of the tcgen05 GEMM shapes of the cfg2 step, 1-CTA vs CTA-pair kernels (CUDA events).
    python scratch/bench_gemm.py"""
import os
import sys
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import _lib, ops  # noqa: E402
from canonicalsg2im_b200.ops import lib    # noqa: E402


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


def rnd(shape, scale=1.0):
    return (torch.randn(shape, device="cuda") * scale).to(torch.bfloat16)


if __name__ == "__main__":
    _lib.load()
    NT, NO = 117321, 2349
    hid = rnd((NT, 512)); w2 = rnd((1152, 512), 0.05); b2 = torch.randn(1152, device="cuda"); conf = torch.rand(NT, device="cuda")
    g = rnd((NT, 1152)); w2t = rnd((512, 1152), 0.05); w1t = rnd((384, 512), 0.05); w1 = rnd((512, 384), 0.05)
    obj, pred = rnd((NO, 128)), rnd((NT, 128))
    s = torch.randint(0, NO, (NT,), device="cuda", dtype=torch.int32); o = torch.randint(0, NO, (NT,), device="cuda", dtype=torch.int32)
    ga = ops.Gather(obj, pred, s, o)
    out2 = torch.empty((NT, 1152), dtype=torch.bfloat16, device="cuda")
    outh = torch.empty((NT, 512), dtype=torch.bfloat16, device="cuda")
    outx = torch.empty((NT, 384), dtype=torch.bfloat16, device="cuda")
    cases = [
        ("F1  gatherA 512x384", 2.0 * NT * 512 * 384, lambda: ops.gemm_bf16(NT, 512, 384, None, w1, bias=b2[:512], relu=True, gather=ga, gather_mode=1, out=outh)),
        ("F2  1152x512 epi", 2.0 * NT * 1152 * 512, lambda: ops.gemm_bf16(NT, 1152, 512, hid, w2, bias=b2, relu=True, rowscale=conf, out=out2)),
        ("dhid 512x1152 mask", 2.0 * NT * 512 * 1152, lambda: ops.gemm_bf16(NT, 512, 1152, g, w2t, mask_aux=hid, out=outh)),
        ("dX  384x512", 2.0 * NT * 384 * 512, lambda: ops.gemm_bf16(NT, 384, 512, hid, w1t, out=outx)),
    ]
    for dbg in [int(x) for x in os.environ.get("CSG_BENCH_DBG", "0,1,2,4,3,5,6").split(",")]:     # 1 = no loads, 2 = no MMAs, 4 = no epilogue (kernel-side switches, wrong results)
        os.environ["CSG_GEMM_DEBUG"] = str(dbg)
        for mode in [int(x) for x in os.environ.get("CSG_BENCH_PAIR", "0,1").split(",")]:
            lib().csg_gemm_bf16_set_pair_mode(mode)
            for name, fl, fn in cases:
                if dbg and name.startswith("F1"):
                    continue
                t = timeit(fn)
                print("dbg=%d pair=%d %-22s %7.1f us  %7.1f TFLOP/s" % (dbg, mode, name, t * 1e6, fl / t / 1e12), flush=True)

# gathered-B weight gradient dW1 = dhid^T [obj[s] | pred | obj[o]]  (M = 512, N = 384, K = triples)
if os.environ.get("CSG_BENCH_DW1"):
    os.environ["CSG_GEMM_DEBUG"] = "0"
    dh = rnd((NT, 512), 0.1)
    fn = lambda: ops.gemm_bf16(512, 384, NT, dh, None, mn_major=True, gather=ga, gather_mode=2)
    t = timeit(fn)
    print("dW1 gatherB 512x384xNT  %7.1f us  %7.1f TFLOP/s" % (t * 1e6, 2.0 * 512 * 384 * NT / t / 1e12), flush=True)
