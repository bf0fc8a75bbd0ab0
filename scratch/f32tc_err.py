"""Accuracy of the fp32 GEMM on the tensor cores (three-term bf16 split, K-concatenated) against the SIMT fp32 kernel,
both measured against float64.   python scratch/f32tc_err.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import ops, _lib
from canonicalsg2im_b200.ops import A_ROW, A_COL, B_NK, B_KN
_lib.load()
def r(shape, seed, scale=1.0):
    g = torch.Generator("cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g) * scale
for M, N, K in [(300, 512, 384), (1000, 1152, 512), (2349, 512, 128), (117321, 512, 1152)]:
    A, B = r((M, K), 1), r((N, K), 2, 0.05)
    ref = A.double() @ B.double().T
    for eng in ("simt", "tc"):
        ops.set_f32_engine(eng)
        out = ops.gemm_f32(A_ROW, B_NK, M, N, K, A, B).double()
        e = (out - ref).abs()
        print("K-major %s M=%d N=%d K=%d  max/max %.2e  rms/rms %.2e  mean signed/rms %.2e" %
              (eng, M, N, K, e.max().item() / ref.abs().max().item(), (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item(),
               ((out - ref).mean() / ref.pow(2).mean().sqrt()).item()), flush=True)
    t0 = torch.matmul(A, B.T).double()
    e = (t0 - ref).abs()
    print("torch fp32 matmul                      max/max %.2e  rms/rms %.2e" % (e.max().item() / ref.abs().max().item(), (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()))
M, N, K = 512, 384, 117321
A, B = r((K, M), 3, 0.1), r((K, N), 4)
ref = A.double().T @ B.double()
for eng in ("simt", "tc"):
    ops.set_f32_engine(eng)
    out = ops.gemm_f32(A_COL, B_KN, M, N, K, A, B).double()
    e = (out - ref).abs()
    print("MN-major %s M=%d N=%d K=%d  max/max %.2e  rms/rms %.2e" % (eng, M, N, K, e.max().item() / ref.abs().max().item(), (e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()), flush=True)
