"""tcgen05 bf16 GEMM family (csrc/gemm_tc.cu) vs torch fp32 matmul on the same bf16-rounded operands.
Tolerance: the inputs are identical bf16 values and accumulation is fp32, so only the summation order and
the bf16 rounding of the output differ: 1e-2 relative (north_star's bf16 MLP budget), 2e-5 for fp32 outputs."""
import pytest
import torch

from tests.util import assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1], ids=["cta1", "cta_pair"], autouse=True)
def pair_mode(request):
    """Every test runs on the 1-CTA kernels and on the CTA-pair (tcgen05 cta_group::2) kernels; the automatic
    policy (pairs once M >= 2 * 128 * #SMs) is restored afterwards and exercised by test_pair_auto_large."""
    from canonicalsg2im_b200.ops import lib
    lib().csg_gemm_bf16_set_pair_mode(request.param)
    yield request.param
    lib().csg_gemm_bf16_set_pair_mode(-1)


def _rand(shape, seed, scale=1.0):
    g = torch.Generator("cuda").manual_seed(seed)
    return (torch.randn(shape, device="cuda", generator=g) * scale).to(torch.bfloat16)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (300, 512, 384), (1000, 1152, 512), (77, 128, 512),
                                   (4096, 384, 512), (513, 192, 1152), (256, 64, 128)])
def test_kmajor_plain(M, N, K):
    from canonicalsg2im_b200 import ops
    A, B = _rand((M, K), 1), _rand((N, K), 2, 0.05)
    ref = A.float() @ B.float().T
    out = ops.gemm_bf16(M, N, K, A, B, out_f32=True)
    assert_close(out, ref, 2e-5, "fp32 out")
    out = ops.gemm_bf16(M, N, K, A, B)
    assert out.dtype == torch.bfloat16
    assert_close(out.float(), ref, 1e-2, "bf16 out")


def test_kmajor_epilogues():
    from canonicalsg2im_b200 import ops
    M, N, K = 777, 512, 384
    A, B = _rand((M, K), 3), _rand((N, K), 4, 0.05)
    bias = torch.randn(N, device="cuda")
    rs = torch.rand(M, device="cuda")
    aux = _rand((M, N), 5)
    z = A.float() @ B.float().T + bias
    assert_close(ops.gemm_bf16(M, N, K, A, B, out_f32=True, bias=bias, relu=True), torch.relu(z), 2e-5, "bias+relu")
    assert_close(ops.gemm_bf16(M, N, K, A, B, out_f32=True, bias=bias, relu=True, rowscale=rs),
                 torch.relu(z) * rs[:, None], 2e-5, "rowscale")
    assert_close(ops.gemm_bf16(M, N, K, A, B, out_f32=True, mask_aux=aux),
                 (A.float() @ B.float().T) * (aux.float() > 0), 2e-5, "mask")
    # strided A (a column slice), as when pred vecs alias the previous layer's output
    big = _rand((M, K + 128), 6)
    assert_close(ops.gemm_bf16(M, N, K, big[:, 64:64 + K], B, out_f32=True), big[:, 64:64 + K].float() @ B.float().T,
                 2e-5, "strided A")


@pytest.mark.parametrize("M,N,K", [(777, 512, 384), (5000, 1152, 512), (40, 128, 512), (1234, 384, 512), (300, 96, 64)])
def test_kmajor_epilogues_bf16_out(M, N, K):
    """bf16 output path: registers -> swizzled smem -> TMA store, ReLU-mask operand prefetched by TMA."""
    from canonicalsg2im_b200 import ops
    A, B = _rand((M, K), 3), _rand((N, K), 4, 0.05)
    bias = torch.randn(N, device="cuda")
    rs = torch.rand(M, device="cuda")
    aux = _rand((M, N), 5)
    z = A.float() @ B.float().T
    out = ops.gemm_bf16(M, N, K, A, B, bias=bias, relu=True, rowscale=rs)
    assert_close(out.float(), torch.relu(z + bias) * rs[:, None], 1e-2, "bias+relu+rowscale bf16")
    out = ops.gemm_bf16(M, N, K, A, B, mask_aux=aux)
    ref = z * (aux.float() > 0)
    assert_close(out.float(), ref, 1e-2, "mask bf16")
    # the mask must be exact: masked entries are exact zeros, kept entries non-zero wherever the reference is
    assert bool(((out.float() == 0) | (aux.float() > 0)).all())
    # output into a column slice of a wider buffer (as new_p aliases the net1 output) + strided aux
    wide = torch.zeros((M, N + 128), dtype=torch.bfloat16, device="cuda")
    auxw = torch.zeros((M, N + 64), dtype=torch.bfloat16, device="cuda")
    auxw[:, 64:] = aux
    ops.gemm_bf16(M, N, K, A, B, out=wide[:, 64:64 + N], mask_aux=auxw[:, 64:])
    assert_close(wide[:, 64:64 + N].float(), ref, 1e-2, "strided C / aux")
    assert float(wide[:, :64].abs().max()) == 0 and float(wide[:, 64 + N:].abs().max()) == 0


def _gather(NO, NT, Din, Dp, seed):
    from canonicalsg2im_b200 import ops
    g = torch.Generator("cuda").manual_seed(seed)
    obj, pred = _rand((NO, Din), seed + 1), _rand((NT, Dp), seed + 2)
    s = torch.randint(0, NO, (NT,), device="cuda", generator=g, dtype=torch.int32)
    o = torch.randint(0, NO, (NT,), device="cuda", generator=g, dtype=torch.int32)
    X = torch.cat([obj[s.long()], pred, obj[o.long()]], 1)
    return ops.Gather(obj, pred, s, o), X


@pytest.mark.parametrize("NT", [1, 130, 5000])
def test_gather_a(NT):
    from canonicalsg2im_b200 import ops
    g, X = _gather(97, NT, 128, 128, 10)
    W = _rand((512, 384), 11, 0.05)
    bias = torch.randn(512, device="cuda")
    ref = torch.relu(X.float() @ W.float().T + bias)
    out = ops.gemm_bf16(NT, 512, 384, None, W, out_f32=True, bias=bias, relu=True, gather=g, gather_mode=1)
    assert_close(out, ref, 2e-5, "gather A")
    out = ops.gemm_bf16(NT, 512, 384, None, W, bias=bias, relu=True, gather=g, gather_mode=1)
    assert_close(out.float(), ref, 1e-2, "gather A, bf16 out")


def test_gather_a_strided_pred():
    """pred rows that alias a wider buffer (the previous layer's [NT, 2H+Dp] output, graph.py:79-81)."""
    from canonicalsg2im_b200 import ops
    g, X = _gather(97, 3000, 128, 128, 12)
    wide = _rand((3000, 1152), 13)
    wide[:, 512:640] = g.pred
    g2 = ops.Gather(g.obj, wide[:, 512:640], g.s_idx, g.o_idx)
    W = _rand((512, 384), 11, 0.05)
    ref = X.float() @ W.float().T
    out = ops.gemm_bf16(3000, 512, 384, None, W, out_f32=True, gather=g2, gather_mode=1)
    assert_close(out, ref, 2e-5, "gather A, strided pred")


@pytest.mark.parametrize("M,N,K", [(512, 512, 64), (1152, 512, 3000), (128, 512, 200), (512, 384, 10000)])
def test_mnmajor(M, N, K):
    from canonicalsg2im_b200 import ops
    A, B = _rand((K, M), 20, 0.1), _rand((K, N), 21, 0.1)
    ref = A.float().T @ B.float()
    assert_close(ops.gemm_bf16(M, N, K, A, B, mn_major=True), ref, 2e-5, "MN-major")


@pytest.mark.parametrize("NT", [64, 1000, 20000])
def test_gather_b(NT):
    from canonicalsg2im_b200 import ops
    g, X = _gather(300, NT, 128, 128, 30)
    dh = _rand((NT, 512), 31, 0.1)
    ref = dh.float().T @ X.float()
    out = ops.gemm_bf16(512, 384, NT, dh, None, mn_major=True, gather=g, gather_mode=2)
    assert_close(out, ref, 2e-5, "gather B")


def test_pair_auto_large(pair_mode):
    """Bench-sized K-major GEMMs (M = 117 321 triples: odd number of 128-row tiles, so the last pair has an empty
    peer CTA) with the automatic pair policy: F2-type epilogue and the masked dX-type, many tiles per pair."""
    if pair_mode == 0:
        pytest.skip("automatic policy runs once")
    from canonicalsg2im_b200 import ops
    from canonicalsg2im_b200.ops import lib
    lib().csg_gemm_bf16_set_pair_mode(-1)
    M = 117321
    A, B = _rand((M, 512), 40), _rand((1152, 512), 41, 0.05)
    bias = torch.randn(1152, device="cuda")
    rs = torch.rand(M, device="cuda")
    out = ops.gemm_bf16(M, 1152, 512, A, B, bias=bias, relu=True, rowscale=rs)
    ref = torch.relu(A.float() @ B.float().T + bias) * rs[:, None]
    assert_close(out.float(), ref, 1e-2, "F2 shape, pairs")
    del ref
    aux = _rand((M, 512), 42)
    B2 = _rand((512, 1152), 43, 0.05)
    out2 = ops.gemm_bf16(M, 512, 1152, out, B2, mask_aux=aux)
    ref2 = (out.float() @ B2.float().T) * (aux.float() > 0)
    assert_close(out2.float(), ref2, 1e-2, "dhid shape, pairs")
    g, X = _gather(2349, M, 128, 128, 44)
    W = _rand((512, 384), 45, 0.05)
    out3 = ops.gemm_bf16(M, 512, 384, None, W, bias=bias[:512].contiguous(), relu=True, gather=g, gather_mode=1)
    assert_close(out3.float(), torch.relu(X.float() @ W.float().T + bias[:512]), 1e-2, "F1 shape, pairs")


@pytest.mark.parametrize("M,N,K", [(128 * 326, 512, 128), (256 * 90, 512, 192), (128 * 190, 128, 256), (117321, 512, 64)])
def test_tail_split_of_the_tile_schedule(M, N, K, monkeypatch):
    """When the last round of the static tile schedule would leave more than half of the workers idle, its tiles are
    cut into two half-width tiles (csrc/gemm_tc.cu, TcParams::tail_from).  Shapes chosen so that the split triggers on
    the 1-CTA and / or the CTA-pair kernels: the result must be bit-identical to the unsplit schedule
    (CSG_GEMM_TAIL_SPLIT=0) -- same products, same K order per output element -- with every epilogue."""
    from canonicalsg2im_b200 import ops
    A, B = _rand((M, K), 11), _rand((N, K), 12, 0.05)
    bias = torch.randn(N, device="cuda")
    rs = torch.rand(M, device="cuda")
    aux = _rand((M, N), 13)
    outs = []
    for split in ("1", "0"):
        monkeypatch.setenv("CSG_GEMM_TAIL_SPLIT", split)
        outs.append((ops.gemm_bf16(M, N, K, A, B, bias=bias, relu=True, rowscale=rs), ops.gemm_bf16(M, N, K, A, B, mask_aux=aux)))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])
    ref = torch.relu(A[-300:].float() @ B.float().T + bias) * rs[-300:, None]
    assert_close(outs[0][0][-300:].float(), ref, 1e-2, "tail rows")
