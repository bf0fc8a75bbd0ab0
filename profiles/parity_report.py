#!/usr/bin/env python
"""Parity table of the CUDA engines against the CPU oracle on the GPU box (not a test: prints / stores every error so
that the tolerances written in tests/ can be checked against what the kernels actually achieve).

    python profiles/parity_report.py [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np   # noqa: E402
import torch         # noqa: E402

from tests import baseline_cases as bc, parity, golden_inputs as gi   # noqa: E402
from tests.util import t   # noqa: E402


def layer_report():
    """tests/golden/gconv_layer.npz inputs: bf16 layer vs the oracle layer (fp32) in both metrics."""
    from canonicalsg2im_b200.graph import GraphTripleConv
    from oracle import graph as ograph
    obj, pred, s, o, p, ty = gi.layer_inputs()
    st = gi.layer_state()
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    rs = {k: T(v).clone().requires_grad_(True) for k, v in st.items()}
    oo, pp = T(obj).clone().requires_grad_(True), T(pred).clone().requires_grad_(True)
    edges = torch.stack([T(s), T(o)], -1)
    ro, rp = ograph.graph_triple_conv(rs, "", oo, pp, edges, T(p) != 0, T(ty), T(p), rs["predicates_transitive_weights"], 512, 128)
    go, gp = gi.layer_out_grads(ro.shape, rp.shape)
    ((ro * T(go)).sum() + (rp * T(gp)).sum()).backward()
    out = {}
    for prec in ("fp32", "bf16"):
        w = torch.nn.Parameter(t(st["predicates_transitive_weights"]))
        layer = GraphTripleConv(128, 128, 128, 128, 512, 1, predicates_transitive_weights=w, precision=prec).cuda()
        layer.load_state_dict({k: t(v) for k, v in st.items()}, strict=True)
        co, cp = t(obj).requires_grad_(True), t(pred).requires_grad_(True)
        a, b = layer(co, cp, edges.cuda(), t(p) != 0, t(ty), t(p))
        ((a.float() * t(go)).sum() + (b.float() * t(gp)).sum()).backward()
        tab = {"new_obj": parity.errs(a.float(), ro), "new_p": parity.errs(b.float(), rp),
               "d_obj": parity.errs(co.grad, oo.grad), "d_pred": parity.errs(cp.grad, pp.grad)}
        for name, prm in layer.named_parameters():
            tab["d " + name] = parity.errs(prm.grad, rs[name].grad)
        out[prec] = tab
    return out


def model_fixture_report():
    """tests/golden/sg2layout_model.npz (B = 4, <= 8 objects): both engines vs the reference golden."""
    import argparse
    from canonicalsg2im_b200 import synth
    from canonicalsg2im_b200.model import Sg2LayoutModel
    g = np.load(os.path.join(ROOT, "tests", "golden", "sg2layout_model.npz"))
    vocab = synth.Vocab(0)
    opt = argparse.Namespace(
        vocab={"attributes": {"objects": {str(i): i for i in range(vocab.num_obj_classes)}},
               "pred_idx_to_name": vocab.pred_names, "pred_name_to_idx": vocab.pred_ids},
        embedding_dim=128, gconv_dim=128, gconv_hidden_dim=512, gconv_pooling="avg", gconv_num_layers=5,
        mlp_normalization="none", mask_size=0, learned_init="uniform")
    out = {}
    for prec in ("fp32", "bf16"):
        model = Sg2LayoutModel(opt, precision=prec).cuda()
        st = {k: t(v) for k, v in gi.model_state().items()}
        for i in range(5):
            st["gconvs.%d.predicates_transitive_weights" % i] = st["trans_candidates_weights"]
        model.load_state_dict(st, strict=True)
        obj_vecs, boxes, _ = model(t(g["objs"]), t(g["triplets"]), t(g["types"]))
        loss = boxes.float().pow(2).sum() + (obj_vecs.float() * t(gi.model_obj_grad(obj_vecs.shape))).sum()
        loss.backward()
        tab = {"obj_vecs": parity.errs(obj_vecs.float(), g["obj_vecs"]), "boxes_pred": parity.errs(boxes.float(), g["boxes_pred"]),
               "loss": {"max": abs(loss.item() - float(g["loss"])) / abs(float(g["loss"])), "l2": 0.0}}
        for name, prm in model.named_parameters():
            if prm.grad is None:
                continue
            if "d_" + name in g.files:
                tab["d " + name] = parity.errs(prm.grad, g["d_" + name])
            elif "dsub_" + name in g.files:
                tab["d " + name] = parity.errs(prm.grad.reshape(-1)[::gi.GRAD_STRIDE], g["dsub_" + name])
            if "dnorm_" + name in g.files and "d " + name in tab:
                nrm = float(g["dnorm_" + name])
                tab["d " + name]["norm"] = abs(prm.grad.double().norm().item() - nrm) / nrm
        out[prec] = tab
    return out


def main():
    rep = {"layer_fixture": layer_report(), "model_fixture": model_fixture_report()}
    for name, case in (("cfg2", bc.cfg2_case()), ("cfg4", bc.cfg4_case())):
        ref = parity.oracle_run(case)
        ref64 = parity.oracle_run(case, torch.float64)
        rep["%s_reference_fp32_vs_f64" % name] = parity.compare(ref64, ref)
        for prec in ("fp32", "bf16"):
            try:
                got = parity.cuda_run(case, ref, prec)
                rep["%s_%s" % (name, prec)] = parity.compare(ref, got)
                rep["%s_%s_vs_f64" % (name, prec)] = parity.compare(ref64, got)
            except Exception as ex:
                print("FAILED", name, prec, repr(ex)[:300])
    for k, tab in rep.items():
        if k in ("layer_fixture", "model_fixture"):
            for prec, tb in tab.items():
                worst = max(tb.items(), key=lambda kv: kv[1]["l2"])
                print("%s %-5s worst l2 %-28s %.3e   worst max %.3e" % (k, prec, worst[0], worst[1]["l2"],
                                                                       max(v["max"] for v in tb.values())))
            continue
        acts = {n: v for n, v in tab.items() if not n.startswith("d ")}
        grads = {n: v for n, v in tab.items() if n.startswith("d ")}
        wl2 = max(grads.items(), key=lambda kv: kv[1]["l2"])
        wmx = max(grads.items(), key=lambda kv: kv[1]["max"])
        print("%-10s fwd max %.3e | grads: worst l2 %.3e (%s), worst max %.3e (%s)" % (
            k, max(v["max"] for v in acts.values()), wl2[1]["l2"], wl2[0], wmx[1]["max"], wmx[0]))
    path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.json")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    json.dump(rep, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
