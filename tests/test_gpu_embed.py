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
    """AttributeEmbeddings with several attributes + attribute_fc_gen (CLEVR layout, attribute_embed.py:18-48)."""
    from canonicalsg2im_b200.model import AttributeEmbeddings
    attrs = {"shape": {"a": 0, "b": 1, "c": 2, "d": 3}, "color": {str(i): i for i in range(9)},
             "size": {"s": 0, "l": 1, "x": 2}}
    torch.manual_seed(0)
    m = AttributeEmbeddings(attrs, 32).cuda()
    x = torch.stack([torch.randint(0, 4, (6, 5)), torch.randint(0, 9, (6, 5)), torch.randint(0, 3, (6, 5))], -1).cuda()
    y = m(x)
    ref = torch.cat([F.embedding(x[..., k], m._modules["att_emb_%d" % k].weight) for k in range(3)], -1)
    ref = m.attribute_fc_gen(ref)
    assert y.shape == (6, 5, 96)
    assert_close(y, ref, 1e-6, "attribute embeddings")


@pytest.mark.parametrize("n", [1, 9, 2349])
def test_masked_box_loss(n):
    from canonicalsg2im_b200.model import masked_box_loss
    rng = np.random.RandomState(n)
    gt = rng.rand(n, 4).astype(np.float32)
    gt[::7] = -1.0                                     # __image__ dummies
    pred = (gt + rng.randn(n, 4).astype(np.float32) * 1.5).astype(np.float32)
    p = t(pred).requires_grad_(True)
    loss = masked_box_loss(p, t(gt))
    (3.0 * loss).backward()
    pc = torch.from_numpy(pred).requires_grad_(True)
    g = torch.from_numpy(gt)
    real = (g >= 0).all(-1)
    if real.any():
        ref = F.smooth_l1_loss(pc[real], g[real])
        (3.0 * ref).backward()
        assert_close(loss, ref, 1e-5, "loss")
        assert_close(p.grad, pc.grad, 1e-5, "dpred")
    else:
        assert loss.item() == 0.0 and (p.grad == 0).all()
