"""Run one BASELINE-shaped model case (tests/baseline_cases.py) through the CPU oracle and through the CUDA path and
tabulate the differences.  Used by tests/test_gpu_baseline_shapes.py (asserts) and profiles/parity_report.py (prints).

Metrics (both relative to the scale of the reference tensor):
  max  = max |a - b| / max |b|        the max-norm form used for activations and for the fp32 contract (1e-5)
  l2   = ||a - b||_2 / ||b||_2        relative L2, used for the weight gradients of the bf16 engine (1e-2)
"""
import numpy as np
import torch

from tests import baseline_cases as bc


def errs(a, b):
    a = a.detach().double().cpu().reshape(-1)
    b = (b.detach() if torch.is_tensor(b) else torch.from_numpy(np.asarray(b))).double().cpu().reshape(-1)
    d = a - b
    return {"max": d.abs().max().item() / max(b.abs().max().item(), 1e-30),
            "l2": d.norm().item() / max(b.norm().item(), 1e-30)}


def oracle_run(case, dtype=torch.float32):
    """Canonicalize with the oracle (pinned to the reference's add_learnt_triplets by make_golden), pad like the
    reference collate, run the oracle model fwd + (box loss + linear functional) bwd.  Returns inputs and results.
    ``dtype=torch.float64`` evaluates the same graph in double precision: the yardstick that separates kernel error
    from the fp32 rounding noise of the reference itself (see tests/test_gpu_baseline_shapes.py)."""
    from oracle import canon as ocanon, graph as ograph, step as ostep
    vocab, graphs, W, seeds, st, opt = case
    canon = []
    for g, seed in zip(graphs, seeds):
        tr, _, ty, _ = ocanon.add_learnt_triplets(g.triplets, vocab.num_preds, vocab.meta_ids, W, True, True,
                                                  bc.canon_uniforms(g, seed))
        canon.append((tr, ty))
    objs, boxes, trips, types = bc.pad_batch(vocab, graphs, canon)
    state = {k: torch.from_numpy(v).to(dtype).clone().requires_grad_(k != "converse_candidates_weights")
             for k, v in st.items()}
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    obj_vecs, boxes_pred = ograph.sg2layout_forward(state, T(objs), T(trips), T(types), vocab.padding_id)
    bl, _ = ostep.bbox_pred_loss(boxes_pred, T(boxes).to(dtype), T(objs))
    loss = bl + (obj_vecs * T(bc.obj_grad(obj_vecs.shape)).to(dtype)).sum() * 1e-2
    loss.backward()
    grads = {k: v.grad for k, v in state.items() if v.grad is not None}
    return dict(objs=objs, boxes=boxes, trips=trips, types=types, canon=canon, obj_vecs=obj_vecs.detach(),
                boxes_pred=boxes_pred.detach(), loss=loss.item(), grads=grads)


def cuda_forward(case, ref, precision):
    """Forward only, under torch.no_grad() (what the inference-only fp16 precision allows)."""
    from canonicalsg2im_b200.model import Sg2LayoutModel
    vocab, graphs, W, seeds, st, opt = case
    model = Sg2LayoutModel(opt, precision=precision).cuda()
    sd = {k: torch.from_numpy(v) for k, v in st.items()}
    for i in range(len(model.gconvs)):
        sd["gconvs.%d.predicates_transitive_weights" % i] = sd["trans_candidates_weights"]
    model.load_state_dict(sd, strict=True)
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    with torch.no_grad():
        obj_vecs, boxes_pred, _ = model(T(ref["objs"]), T(ref["trips"]), T(ref["types"]))
    return dict(obj_vecs=obj_vecs.float(), boxes_pred=boxes_pred.float())


def cuda_run(case, ref, precision):
    """The same padded batch through canonicalsg2im_b200.Sg2LayoutModel on cuda (reference state-dict keys)."""
    from canonicalsg2im_b200.model import Sg2LayoutModel, bbox_pred_loss
    vocab, graphs, W, seeds, st, opt = case
    model = Sg2LayoutModel(opt, precision=precision).cuda()
    sd = {k: torch.from_numpy(v) for k, v in st.items()}
    for i in range(len(model.gconvs)):
        sd["gconvs.%d.predicates_transitive_weights" % i] = sd["trans_candidates_weights"]
    model.load_state_dict(sd, strict=True)
    T = lambda x: torch.from_numpy(np.ascontiguousarray(x)).cuda()
    obj_vecs, boxes_pred, _ = model(T(ref["objs"]), T(ref["trips"]), T(ref["types"]))
    bl, _ = bbox_pred_loss(boxes_pred.float(), T(ref["boxes"]), T(ref["objs"]))
    loss = bl + (obj_vecs.float() * T(bc.obj_grad(obj_vecs.shape))).sum() * 1e-2
    loss.backward()
    grads = {n: p.grad for n, p in model.named_parameters()
             if p.grad is not None and "predicates_transitive_weights" not in n}
    return dict(obj_vecs=obj_vecs.detach().float(), boxes_pred=boxes_pred.detach().float(), loss=loss.item(), grads=grads)


def compare(ref, got):
    table = {"obj_vecs": errs(got["obj_vecs"], ref["obj_vecs"]), "boxes_pred": errs(got["boxes_pred"], ref["boxes_pred"]),
             "loss": {"max": abs(got["loss"] - ref["loss"]) / abs(ref["loss"]), "l2": abs(got["loss"] - ref["loss"]) / abs(ref["loss"])}}
    for k, g in ref["grads"].items():
        assert k in got["grads"], "missing gradient " + k
        table["d " + k] = errs(got["grads"][k], g)
    return table


def check_oracle_against_golden(ref, g, stride):
    """The oracle run on the box reproduces the committed outputs of the unmodified reference (pins the oracle here)."""
    out = {"obj_vecs": errs(ref["obj_vecs"], g["obj_vecs"])["max"], "boxes_pred": errs(ref["boxes_pred"], g["boxes_pred"])["max"],
           "loss": abs(ref["loss"] - float(g["loss"])) / abs(float(g["loss"]))}
    assert [len(tr) for tr, _ in ref["canon"]] == list(g["n_canon"]) and int(ref["types"].sum()) == int(g["types_sum"])
    for k, gr in ref["grads"].items():
        if "d_" + k in g.files:
            out["d " + k] = errs(gr, g["d_" + k])["max"]
        elif "dsub_" + k in g.files:
            out["d " + k] = errs(gr.reshape(-1)[::stride], g["dsub_" + k])["max"]
    return out
