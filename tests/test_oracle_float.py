"""Oracle (oracle/graph.py, oracle/layout.py) against golden outputs of the reference — CPU only."""
import numpy as np
import torch

from oracle import graph as ograph, layout as olayout
from tests import golden_inputs as gi


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def test_graph_triple_conv_fwd_bwd(golden):
    g = golden("gconv_layer")
    obj, pred, s, o, p, ty = gi.layer_inputs()
    st = {k: t(v).clone().requires_grad_(True) for k, v in gi.layer_state().items()}
    oo, pp = t(obj).clone().requires_grad_(True), t(pred).clone().requires_grad_(True)
    edges = torch.stack([t(s), t(o)], -1)
    new_obj, new_p = ograph.graph_triple_conv(st, "", oo, pp, edges, t(p) != 0, t(ty), t(p),
                                              st["predicates_transitive_weights"], 512, 128)
    go, gp = gi.layer_out_grads(new_obj.shape, new_p.shape)
    ((new_obj * t(go)).sum() + (new_p * t(gp)).sum()).backward()
    # tolerance: fp32, 1e-5 relative to the tensor scale (north_star)
    def close(a, b, tol=1e-5):
        b = t(b)
        assert (a.detach() - b).abs().max() <= tol * max(1.0, b.abs().max().item())
    close(new_obj, g["new_obj"]); close(new_p, g["new_p"])
    close(oo.grad, g["d_obj"]); close(pp.grad, g["d_pred"])
    close(st["predicates_transitive_weights"].grad, g["d_w_trans"])
    for k in ["net1.0.weight", "net1.2.bias", "net2.2.weight"]:
        close(st[k].grad.reshape(-1)[::gi.GRAD_STRIDE], g["dsub_" + k])
    # edge types 2/3 give exactly-zero predicate rows
    assert (new_p.detach()[t(ty) >= 2] == 0).all()


def test_sg2layout_model(golden):
    g = golden("sg2layout_model")
    st = {k: t(v) for k, v in gi.model_state().items()}
    vecs, boxes = ograph.sg2layout_forward(st, t(g["objs"]), t(g["triplets"]), t(g["types"]), 0)
    assert torch.allclose(vecs, t(g["obj_vecs"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(boxes, t(g["boxes_pred"]), rtol=1e-5, atol=1e-5)


def test_layout_cases(golden):
    g = golden("layout")
    v, b, m = t(g["demo_vecs"]), t(g["demo_boxes"]), t(g["demo_masks"])
    assert torch.equal(olayout.boxes_to_layout(v, b, 64), t(g["demo_boxes64_out"]))
    assert torch.equal(olayout.masks_to_layout(v, b, m, 64), t(g["demo_masks64_out"]))
    assert torch.equal(olayout.masks_to_layout(v, b, m, 64, test_mode=True), t(g["demo_masks64_test_out"]))
    assert torch.equal(olayout.boxes_to_layout(v, b, 64, align_corners=True), t(g["demo_boxes64_legacy_out"]))
    rv, rb, rm = t(g["rnd_vecs"]), t(g["rnd_boxes"]), t(g["rnd_masks"])
    assert torch.equal(olayout.boxes_to_layout(rv, rb, 32, 48), t(g["rnd_boxes_out"]))
    assert torch.equal(olayout.masks_to_layout(rv, rb, rm, 32, 48), t(g["rnd_masks_out"]))
    assert torch.equal(olayout.masks_to_layout(rv, rb, rm, 32, 48, test_mode=True), t(g["rnd_masks_test_out"]))
    assert torch.allclose(olayout.boxes_to_layout(rv, rb, 32, 48, pooling="avg"), t(g["rnd_boxes_avg_out"]))


def test_layout_grads(golden):
    g = golden("layout")
    v = t(g["rnd_vecs"]).clone().requires_grad_(True)
    b = t(g["rnd_boxes"]).clone().requires_grad_(True)
    m = t(g["rnd_masks_f_in"]).clone().requires_grad_(True)
    y = olayout.masks_to_layout(v, b, m, 40, 40)
    (y * t(gi.layout_out_grad(y.shape))).sum().backward()
    assert torch.allclose(v.grad, t(g["rnd_masks_f_dvecs"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(b.grad, t(g["rnd_masks_f_dboxes"]), rtol=1e-4, atol=1e-4)
    assert torch.allclose(m.grad, t(g["rnd_masks_f_dmasks"]), rtol=1e-5, atol=1e-5)


def test_crop_bbox(golden):
    g = golden("layout")
    im = t(g["crop_imgs"]).clone().requires_grad_(True)
    c = olayout.crop_bbox(im, t(g["crop_boxes"]), 8, 12)
    assert torch.equal(c.detach(), t(g["crop_out"]))
    (c * t(gi.crop_out_grad(c.shape))).sum().backward()
    assert torch.allclose(im.grad, t(g["crop_dimgs"]), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------------
# BASELINE-shaped fixtures (round 2): the oracle against outputs of the unmodified reference at cfg2 / cfg4 / cfg3 shapes
# ------------------------------------------------------------------------------------------------
def test_oracle_model_at_cfg2_and_cfg4_shapes(golden):
    from tests import baseline_cases as bc, parity
    for name, case in (("cfg2_model", bc.cfg2_case()), ("cfg4_model", bc.cfg4_case())):
        out = parity.check_oracle_against_golden(parity.oracle_run(case), golden(name), gi.GRAD_STRIDE)
        bad = {k: v for k, v in out.items() if v > 2e-5}
        assert not bad and len(out) >= 47, (name, bad)


def test_oracle_bbox_pred_loss(golden):
    from canonicalsg2im_b200 import synth
    from oracle import step as ostep
    g = golden("box_loss")
    for ci in range(int(g["num_cases"])):
        A, B, O, seed = [int(x) for x in g["c%d_spec" % ci]]
        objs = synth.det_int(B * O * A, seed * 11 + 1, 0 if A > 1 else 1, 7).reshape(B, O, A)
        n_real = synth.det_int(B, seed * 11 + 2, 1, O - 1)
        boxes = synth.det_tensor((B, O, 4), seed * 11 + 3, 0.5) + np.float32(0.5)
        for b in range(B):
            objs[b, n_real[b]:] = 0
            boxes[b, n_real[b]:] = -1.0
            if A > 1:
                objs[b, :n_real[b], 0] = np.maximum(objs[b, :n_real[b], 0], 1)
        pred = t(boxes + synth.det_tensor((B, O, 4), seed * 11 + 4, 2.0)).requires_grad_(True)
        loss, loss_all = ostep.bbox_pred_loss(pred, t(boxes), t(objs))
        (3.0 * loss).backward()
        assert abs(loss.item() - float(g["c%d_loss" % ci])) <= 1e-6 * float(g["c%d_loss" % ci])
        assert torch.allclose(loss_all, t(g["c%d_loss_all" % ci]), rtol=1e-6, atol=0)
        assert torch.allclose(pred.grad, t(g["c%d_dpred" % ci]), rtol=1e-6, atol=1e-9)


def test_oracle_layout_at_cfg3_shape(golden):
    from canonicalsg2im_b200 import synth
    g = golden("cfg3_layout")
    gr = synth.make_graph(777, 8, 8, synth.Vocab(0), include_dummies=False, mask_size=16)
    vecs = t(synth.det_tensor((8, 128), 12, 1.0))
    y = olayout.masks_to_layout(vecs, t(gr.boxes), t(gr.masks), 256, 256)
    yt = olayout.masks_to_layout(vecs, t(gr.boxes), t(gr.masks), 256, 256, test_mode=True)
    assert torch.equal(y[:, :, ::8, ::8], t(g["train_sub"])) and torch.equal(yt[:, :, ::8, ::8], t(g["test_sub"]))
    assert abs(y.double().sum().item() - float(g["train_sum"])) <= 1e-9 * float(g["train_abs"])   # summation order only
