"""Embedding lookups (sg2im/model.py:108-109, attribute_embed.py:38-48) and the masked box loss
(pix2pix_model.py:72-85) on the csg2im kernels vs plain torch fp32 on the CPU.
Tolerance: lookups are copies (bit-exact in fp32); sums 1e-5 relative (fp32), 1e-2 for bf16 inputs."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from canonicalsg2im_b200 import synth
from tests.util import t, assert_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,V,E", [(0, 7, 8), (1, 3, 4), (1000, 50, 128), (2349, 184, 128), (5000, 300, 256)])
def test_embedding_fwd_bwd_fp32(n, V, E):
    from canonicalsg2im_b200.model import embedding_lookup
    rng = np.random.RandomState(n + V)
    table = synth.det_tensor((V, E), 3, 1.0)
    idx = rng.randint(0, V, size=(n,)).astype(np.int64)
    gy = synth.det_tensor((n, E), 4, 1.0) if n else np.zeros((0, E), np.float32)
    w = t(table).requires_grad_(True)
    y = embedding_lookup(w, t(idx))
    assert y.shape == (n, E)
    wc = torch.from_numpy(table).requires_grad_(True)
    yc = F.embedding(torch.from_numpy(idx), wc)
    assert torch.equal(y.detach().cpu(), yc.detach())
    (y * t(gy)).sum().backward()
    (yc * torch.from_numpy(gy)).sum().backward()
    assert_close(w.grad, wc.grad, 1e-5, "dtable")


def test_embedding_strided_index_and_bf16():
    """The index is a strided column of the int64 triplets tensor; bf16 rows out, bf16 gradient rows in
    (a column slice of a wider matrix, as the GCN's dX hands it over)."""
    from canonicalsg2im_b200.model import embedding_lookup
    rng = np.random.RandomState(5)
    V, E, n = 50, 128, 4097
    trip = rng.randint(0, V, size=(n, 3)).astype(np.int64)
    table = synth.det_tensor((V, E), 6, 1.0)
    w = t(table).requires_grad_(True)
    y = embedding_lookup(w, t(trip)[:, 1], torch.bfloat16)
    ref = torch.from_numpy(table)[torch.from_numpy(trip[:, 1])].to(torch.bfloat16)
    assert torch.equal(y.detach().cpu(), ref)
    wide = t(synth.det_tensor((n, 3 * E), 7, 1.0)).to(torch.bfloat16)
    y.backward(wide[:, E:2 * E])
    gref = torch.zeros(V, E).index_add_(0, torch.from_numpy(trip[:, 1]), wide[:, E:2 * E].float().cpu())
    assert_close(w.grad, gref, 1e-5, "dtable from bf16 rows")
    # run-to-run determinism
    w2 = t(table).requires_grad_(True)
    embedding_lookup(w2, t(trip)[:, 1], torch.bfloat16).backward(wide[:, E:2 * E])
    assert torch.equal(w.grad, w2.grad)


def test_embedding_bwd_large_bf16_tensor_core_path():
    """n >= 8192 bf16 gradient rows: dtable = onehot^T dout on the tcgen05 MN-major GEMM (exact products, fp32
    accumulation): 1e-5 against an fp32 index_add of the same bf16 values, deterministic, unused ids exactly zero."""
    from canonicalsg2im_b200.model import embedding_lookup
    rng = np.random.RandomState(9)
    V, E, n = 51, 128, 117321
    trip = rng.randint(0, V - 1, size=(n, 3)).astype(np.int64)       # id V-1 never used
    table = synth.det_tensor((V, E), 6, 1.0)
    wide = t(synth.det_tensor((n, 3 * E), 8, 1.0)).to(torch.bfloat16)
    grads = []
    for _ in range(2):
        w = t(table).requires_grad_(True)
        embedding_lookup(w, t(trip)[:, 1], torch.bfloat16).backward(wide[:, E:2 * E])
        grads.append(w.grad)
    gref = torch.zeros(V, E, dtype=torch.float64).index_add_(0, torch.from_numpy(trip[:, 1]),
                                                             wide[:, E:2 * E].double().cpu()).float()
    assert_close(grads[0], gref, 1e-5, "dtable, tensor-core path")
    assert torch.equal(grads[0], grads[1])
    assert float(grads[0][V - 1].abs().max()) == 0.0


def test_model_embeddings_match_torch_modules():
    """AttributeEmbeddings with several attributes + attribute_fc_gen (CLEVR layout, attribute_embed.py:18-48): lookups
    (csg_embed_fwd writing column slices) and the Linear (csg_gemm_f32 / csg_gemm_bf16, no cuBLAS) against plain torch
    ops on the CPU, forward and every gradient."""
    from canonicalsg2im_b200.model import AttributeEmbeddings
    attrs = {"shape": {"a": 0, "b": 1, "c": 2, "d": 3}, "color": {str(i): i for i in range(9)},
             "size": {"s": 0, "l": 1, "x": 2}, "material": {"r": 0, "m": 1, "q": 2}}
    torch.manual_seed(0)
    for emb, out_dtype, tol in ((32, torch.float32, 1e-5), (24, torch.float32, 1e-5), (32, torch.bfloat16, 1e-2)):
        m = AttributeEmbeddings(attrs, emb).cuda()
        x = torch.stack([torch.randint(0, 4, (6, 50)), torch.randint(0, 9, (6, 50)), torch.randint(0, 3, (6, 50)),
                         torch.randint(0, 3, (6, 50))], -1).cuda()
        y = m(x, out_dtype)
        assert y.shape == (6, 50, 4 * emb) and y.dtype == out_dtype
        gy = torch.randn(y.shape, device="cuda")
        (y.float() * gy).sum().backward()
        cpu = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in m.named_parameters()}
        ref = torch.cat([F.embedding(x[..., k].cpu(), cpu["att_emb_%d.weight" % k]) for k in range(4)], -1)
        ref = F.linear(ref, cpu["attribute_fc_gen.weight"], cpu["attribute_fc_gen.bias"])
        (ref * gy.cpu()).sum().backward()
        assert_close(y.float(), ref, tol, "attribute embeddings")
        for k, v in m.named_parameters():
            a, b_ = v.grad.double().cpu(), cpu[k].grad.double()
            assert ((a - b_).norm() / b_.norm()).item() <= tol, k


@pytest.mark.parametrize("B,O", [(1, 3), (7, 9), (128, 31)])
def test_bbox_pred_loss_vs_torch(B, O):
    """csg_box_loss vs the reference formula (pix2pix_model.py:72-85) written with torch ops on the CPU."""
    from canonicalsg2im_b200.model import bbox_pred_loss
    rng = np.random.RandomState(B)
    objs = rng.randint(1, 100, size=(B, O, 1))
    gt = rng.rand(B, O, 4).astype(np.float32)
    for b in range(B):
        n = rng.randint(1, O)
        objs[b, n:] = 0
        gt[b, n:] = -1.0
    pred = (gt + rng.randn(B, O, 4).astype(np.float32) * 1.5).astype(np.float32)
    p = t(pred).requires_grad_(True)
    loss, loss_all = bbox_pred_loss(p, t(gt), t(objs), weight=10.0)
    (3.0 * loss).backward()
    pc = torch.from_numpy(pred).requires_grad_(True)
    flat = F.smooth_l1_loss(pc.view(-1, 4), torch.from_numpy(gt).view(-1, 4), reduction="none") * 10.0
    real = (torch.from_numpy(objs).view(-1, 1) != 0).float()
    ref_all = (flat * real).view(B, O, 4).sum(dim=[1, 2]) / real.view(B, O).sum(dim=1)
    (3.0 * ref_all.mean()).backward()
    assert_close(loss, ref_all.mean(), 1e-5, "loss")
    assert_close(loss_all, ref_all, 1e-5, "loss_all")
    assert_close(p.grad, pc.grad, 1e-5, "dpred")
