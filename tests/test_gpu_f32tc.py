"""The fp32 parity engine on the tensor cores (ops.F32_ENGINE = "tc"): every GEMM mode of ``ops.gemm_f32`` through an
exact three-term bf16 split of both operands, K-concatenated into tcgen05 GEMMs with fp32 accumulation (csg_split3_bf16
+ csg_gemm_bf16).  tcgen05.mma truncates its accumulator at every instruction, so WHERE the big products sit in the
chain decides the accuracy: with the six products ordered smallest first (hi*hi last) and, for the long reductions of the
weight gradients, hi*hi run as chains of <= ~64 MMAs added in fp32, the engine is as accurate as the fp32 FMA kernel
(measured against float64, relative rms: 1.3e-7 .. 1.1e-6 for K = 128 .. 1152, 4.8e-7 for K = 117 321; FMA kernel:
2.0e-7 .. 6.1e-7 and 1.2e-6; with hi*hi FIRST the same GEMMs were at 1.1e-8 x K, i.e. 1.3e-5 at K = 1152).
Tolerances: 3e-6 max-norm against float64 matmul per GEMM, and for the 5-layer model the SAME golden contract as the
FMA engine (1e-5 outputs / 2e-5 gradients against the unmodified reference's fp32 run)."""
import numpy as np
import pytest
import torch

from canonicalsg2im_b200 import synth
from tests import golden_inputs as gi
from tests.util import t, assert_close

pytestmark = pytest.mark.gpu


@pytest.fixture
def tc_engine():
    from canonicalsg2im_b200 import ops
    old = ops.F32_ENGINE
    ops.set_f32_engine("tc")
    yield ops
    ops.set_f32_engine(old)


def _r(shape, seed, scale=1.0):
    g = torch.Generator("cuda").manual_seed(seed)
    return torch.randn(shape, device="cuda", generator=g) * scale


def test_split3_is_exact():
    """hi + mid + lo reproduces the fp32 value bit for bit (8 + 8 + 8 significant bits), in both layouts."""
    from canonicalsg2im_b200 import ops
    X = _r((70, 50), 0) * torch.logspace(-6, 6, 50, device="cuda")
    for k_is_cols in (True, False):
        for tr in (False, True):
            for role, order in ((0, (1, 2, 0, 1, 0, 0)), (1, (1, 0, 2, 0, 1, 0))):
                out = ops._split3(X, 70, 50, tr, k_is_cols, role).float()
                L = X.T if tr else X
                R, C = L.shape
                parts = [out[:, k * C:(k + 1) * C] if k_is_cols else out[k * R:(k + 1) * R] for k in range(6)]
                first = {order[k]: parts[k] for k in (5, 4, 3, 2, 1, 0)}          # one representative of hi / mid / lo
                assert torch.equal(first[0] + first[1] + first[2], L)
                for k in range(6):
                    assert torch.equal(parts[k], first[order[k]])


@pytest.mark.parametrize("M,N,K", [(300, 512, 384), (1000, 1152, 512), (77, 128, 512), (2349, 512, 128)])
def test_gemm_modes_match_float64(tc_engine, M, N, K):
    ops = tc_engine
    from canonicalsg2im_b200.ops import A_ROW, A_COL, B_NK, B_KN
    A, Bnk, Bkn = _r((M, K), 1), _r((N, K), 2, 0.05), _r((K, N), 3, 0.05)
    bias, rs, mask = _r((N,), 4), _r((M,), 5).abs(), _r((M, N), 6)

    def check(out, ref, what):
        err = (out.double() - ref).abs().max().item() / ref.abs().max().item()
        assert err <= 3e-6, (what, err)

    check(ops.gemm_f32(A_ROW, B_NK, M, N, K, A, Bnk), A.double() @ Bnk.double().T, "A_ROW x B_NK")
    check(ops.gemm_f32(A_ROW, B_KN, M, N, K, A, Bkn), A.double() @ Bkn.double(), "A_ROW x B_KN")
    ref = torch.relu(A.double() @ Bnk.double().T + bias.double()) * rs.double()[:, None]
    check(ops.gemm_f32(A_ROW, B_NK, M, N, K, A, Bnk, bias=bias, relu=True, rowscale=rs), ref, "bias + relu + rowscale")
    ref = (A.double() @ Bkn.double()) * (mask > 0).double()
    out = ops.gemm_f32(A_ROW, B_KN, M, N, K, A, Bkn, mask_aux=mask)
    check(out, ref, "relu mask")
    assert torch.equal(out == 0, ~(mask > 0) | (out == 0))
    # weight-gradient form: C[M', N'] = A[K', M']^T B[K', N'] with the long dimension reduced
    Ak, Bk = _r((M, 512), 7), _r((M, N), 8)
    check(ops.gemm_f32(A_COL, B_KN, 512, N, M, Ak, Bk), Ak.double().T @ Bk.double(), "A_COL x B_KN")


def test_unsupported_shapes_fall_back(tc_engine):
    """Widths the tcgen05 tiles do not take (N % 32 != 0: box_net's 4 outputs) run on the SIMT kernel, same API."""
    ops = tc_engine
    from canonicalsg2im_b200.ops import A_ROW, B_NK
    A, B = _r((100, 512), 1), _r((4, 512), 2)
    out = ops.gemm_f32(A_ROW, B_NK, 100, 4, 512, A, B)
    assert_close(out, (A.double() @ B.double().T).float(), 1e-5, "fallback")


def test_model_golden_on_the_split_engine(tc_engine, golden):
    """tests/test_gpu_graph.py::test_model_golden_fwd_bwd with every GEMM on tcgen05: same tolerances."""
    import argparse
    from canonicalsg2im_b200.model import Sg2LayoutModel
    g = golden("sg2layout_model")
    vocab = synth.Vocab(0)
    opt = argparse.Namespace(
        vocab={"attributes": {"objects": {str(i): i for i in range(vocab.num_obj_classes)}},
               "pred_idx_to_name": vocab.pred_names, "pred_name_to_idx": vocab.pred_ids},
        embedding_dim=128, gconv_dim=128, gconv_hidden_dim=512, gconv_pooling="avg", gconv_num_layers=5,
        mlp_normalization="none", mask_size=0, learned_init="uniform")
    model = Sg2LayoutModel(opt, precision="fp32").cuda()
    st = {k: t(v) for k, v in gi.model_state().items()}
    for i in range(5):
        st["gconvs.%d.predicates_transitive_weights" % i] = st["trans_candidates_weights"]
    model.load_state_dict(st, strict=True)
    obj_vecs, boxes, _ = model(t(g["objs"]), t(g["triplets"]), t(g["types"]))
    assert_close(obj_vecs, g["obj_vecs"], 1e-5, "obj_vecs")
    assert_close(boxes, g["boxes_pred"], 1e-5, "boxes_pred")
    loss = boxes.pow(2).sum() + (obj_vecs * t(gi.model_obj_grad(obj_vecs.shape))).sum()
    loss.backward()
    checked = 0
    for name, prm in model.named_parameters():
        if "d_" + name in g.files:
            assert_close(prm.grad, g["d_" + name], 2e-5, "d " + name); checked += 1
        elif "dsub_" + name in g.files:
            assert_close(prm.grad.reshape(-1)[::gi.GRAD_STRIDE], g["dsub_" + name], 2e-5, "d " + name); checked += 1
    assert checked >= 40
