"""One GEMM shape under a kernel-side debug switch, for ncu: CSG_GEMM_DEBUG=3 python scratch/prof_gemm_epi.py F2"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import _lib, ops
from canonicalsg2im_b200.ops import lib
_lib.load()
lib().csg_gemm_bf16_set_pair_mode(0)
NT = 117321
which = sys.argv[1] if len(sys.argv) > 1 else "F2"
rnd = lambda shape, sc=1.0: (torch.randn(shape, device="cuda") * sc).to(torch.bfloat16)
if which == "F2":
    A, B = rnd((NT, 512)), rnd((1152, 512), 0.05)
    bias, rs = torch.randn(1152, device="cuda"), torch.rand(NT, device="cuda")
    out = torch.empty((NT, 1152), dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.gemm_bf16(NT, 1152, 512, A, B, bias=bias, relu=True, rowscale=rs, out=out)
else:
    A, B = rnd((NT, 512)), rnd((384, 512), 0.05)
    out = torch.empty((NT, 384), dtype=torch.bfloat16, device="cuda")
    fn = lambda: ops.gemm_bf16(NT, 384, 512, A, B, out=out)
for _ in range(5):
    fn()
torch.cuda.synchronize()
