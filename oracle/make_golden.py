"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference.

Run in the authoring container only (``/root/reference`` does not exist on
the GPU box):

    python -m oracle.make_golden            # from the repo root

For every fixture the script (1) builds seeded inputs from
``canonicalsg2im_b200.synth`` (pure integer hashing, reproducible anywhere),
(2) runs the reference's own code on them, (3) asserts that the oracle
restatement under ``oracle/`` reproduces the reference, and (4) stores inputs
that are not regenerable plus the reference outputs.  TEST INFRASTRUCTURE ONLY.
"""
import argparse
import os
import pickle
import sys
import types
import warnings

import numpy as np
import torch

REF = os.environ.get("CSG_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference():
    """sg2im/utils.py:2 imports ``torch.tensor.Tensor`` which no longer exists."""
    shim = types.ModuleType("torch.tensor")
    shim.Tensor = torch.Tensor
    sys.modules.setdefault("torch.tensor", shim)
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import scripts.graphs_utils as gu
    import sg2im.data.base_dataset as bd
    import sg2im.graph as rgraph
    import sg2im.model as rmodel
    import sg2im.layout as rlayout
    import sg2im.bilinear as rbil
    return gu, bd, rgraph, rmodel, rlayout, rbil


sys.path.insert(0, ROOT)
from canonicalsg2im_b200 import synth  # noqa: E402
from oracle import canon as ocanon, graph as ograph, layout as olayout  # noqa: E402
from tests import golden_inputs as gi  # noqa: E402
from tests import baseline_cases as bc  # noqa: E402


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


# ----------------------------------------------------------------------------
def gen_canon(gu, bd):
    out = {}
    # --- reference KAT (graphs_utils.py:159-174) ---
    kat = [[0, 1, 1], [0, 1, 2], [0, 1, 3], [1, 1, 2], [3, 1, 1], [3, 1, 2]]
    adj = np.array(gu.triplets_to_adj_matrix(kat))
    mini = np.array(gu.triplets_to_minimal(kat)).astype(np.int64)
    cur, trans = gu.get_current_and_transitive_triplets(mini)
    cyc_cur, cyc_trans = gu.get_current_and_transitive_triplets(np.array([[0, 1, 1], [1, 1, 0]]))
    out.update(kat_triplets=np.array(kat), kat_adj=adj, kat_minimal=mini,
               kat_transitive=np.array(trans).astype(np.int64),
               cyc_transitive=np.array(cyc_trans).astype(np.int64))
    assert (mini == np.array([[0, 1, 3], [1, 1, 2], [3, 1, 1]])).all()          # graphs_utils.py:166-168
    assert (ocanon.triplets_to_adj(kat) == adj.astype(bool)).all()
    assert (ocanon.triplets_to_minimal(kat) == mini).all()
    assert (ocanon.current_and_transitive(mini)[1] == out["kat_transitive"]).all()
    assert (ocanon.current_and_transitive([[0, 1, 1], [1, 1, 0]])[1] == out["cyc_transitive"]).all()

    # --- random adjacency: closure / hsu / minimal (incl. cyclic graphs) ---
    for i, (n, dens) in enumerate([(5, 0.3), (12, 0.15), (33, 0.08), (64, 0.05), (20, 0.5)]):
        a = (synth.det_uniform(n * n, 900 + i).reshape(n, n) < dens).astype(np.uint8)
        ref_path = np.array(gu.path(a.tolist()), dtype=bool)
        m = [list(r) for r in ref_path.astype(np.uint8)]
        gu.hsu(m)
        ref_min = np.array(m, dtype=bool)
        assert (ocanon.closure(a) == ref_path).all()
        assert (ocanon.hsu_reduce(ocanon.closure(a)) == ref_min).all()
        out["adj%d" % i], out["adj%d_path" % i], out["adj%d_min" % i] = a, ref_path, ref_min
    out["num_adj"] = np.array(5)

    # --- add_location_triplets + add_dummy_triplets + add_learnt_triplets ---
    cases = []
    specs = [  # (vocab, n_min, n_max, include_dummies, box_mode, converse, transitive)
        (synth.Vocab(0), 3, 8, True, "coco", 1, 1),
        (synth.Vocab(0), 3, 8, True, "coco", 0, 0),
        (synth.Vocab(0), 6, 12, False, "coco", 1, 0),
        (synth.Vocab(0), 6, 12, True, "clevr", 0, 1),
        (synth.Vocab(42), 3, 30, True, "coco", 1, 1),
        (synth.Vocab(42), 20, 30, True, "coco", 1, 1),
        (synth.Vocab(0, num_attributes=4), 32, 40, True, "clevr", 1, 1),
    ]
    for ci, (vocab, n0, n1, dummies, mode, conv, trans) in enumerate(specs):
        g = synth.make_graph(4000 + ci, n0, n1, vocab, include_dummies=dummies, box_mode=mode)
        ds = bd.BaseDataset()
        ds.vocab = {"pred_name_to_idx": dict(vocab.pred_ids), "pred_idx_to_name": list(vocab.pred_names),
                    "object_name_to_idx": {"__image__": vocab.image_obj_id},
                    "attributes": {"objects": {"__image__": vocab.image_obj_id}}}
        ds.include_dummies = dummies
        ds.learned_converse, ds.learned_transitivity = bool(conv), bool(trans)
        W = synth.make_conv_weights(vocab, seed=ci)
        ds.converse_candidates_weights = W
        # reference spatial + dummy triples
        n_real = len(g.centers)
        cen = np.concatenate([g.centers, np.zeros((len(g.boxes) - n_real, 2), np.float32)])
        ref_trip = []
        ds.add_location_triplets(t(g.boxes), t(cen), t(g.objs[:, 0]), ref_trip)
        n_loc = len(ref_trip)
        ds.add_dummy_triplets(t(g.objs[:, 0]), ref_trip)
        ref_base = np.array([np.asarray(x) for x in ref_trip]).astype(np.int64)
        # oracle + synth agree with the reference
        o_loc = ocanon.add_location_triplets(g.boxes, cen, g.objs[:, 0], vocab.image_obj_id, vocab.pred_ids)
        assert (np.array(o_loc).reshape(-1, 3) == ref_base[:n_loc]).all(), "oracle add_location_triplets"
        spatial = g.triplets[np.isin(g.triplets[:, 1], [vocab.pred_ids[a] for a in synth.AUGMENTED_RELATIONS])]
        assert (spatial == ref_base[:n_loc]).all(), "synth.location_triplets"
        o_dummy = ocanon.add_dummy_triplets(g.objs[:, 0], vocab.image_obj_id, vocab.in_image_id, dummies)
        assert (np.array(o_dummy).reshape(-1, 3) == ref_base[n_loc:]).all()
        base = g.triplets  # includes the VG-like dataset predicates when present
        seed = 77 + ci
        np.random.seed(seed)
        r_trip, r_counts, r_type = ds.add_learnt_triplets([list(x) for x in base], len(g.objs))
        uniforms = np.random.RandomState(seed).random_sample(len(base) * 2 + 8)
        o_trip, o_counts, o_type, used = ocanon.add_learnt_triplets(
            base, vocab.num_preds, vocab.meta_ids, W, bool(conv), bool(trans), uniforms)
        assert (o_trip == np.asarray(r_trip).astype(np.int64)).all(), "oracle add_learnt_triplets (edges)"
        assert (o_counts == r_counts).all() and (o_type == np.asarray(r_type)).all()
        out["c%d_spec" % ci] = np.array([vocab.num_base_preds, vocab.num_attributes, n0, n1, int(dummies),
                                        int(mode == "clevr"), conv, trans, 4000 + ci, ci, used])
        out["c%d_uniforms" % ci] = uniforms[:max(used, 1)]
        out["c%d_base" % ci] = ref_base
        out["c%d_triplets" % ci] = np.asarray(r_trip).astype(np.int64)
        out["c%d_counts" % ci] = r_counts
        out["c%d_type" % ci] = np.asarray(r_type).astype(np.int64)
        cases.append(ci)
        print("canon case %d: O=%d base=%d -> T'=%d (type1=%d) draws=%d" % (
            ci, len(g.objs), len(base), len(r_trip), int(np.sum(np.asarray(r_type) == 1)), used))
    out["num_cases"] = np.array(len(cases))

    # --- shipped CLEVR graphs (sg2im/data/scene_graphs.pkl), all-pairs relations, closure on ---
    with open(os.path.join(REF, "sg2im", "data", "scene_graphs.pkl"), "rb") as f:
        sgs = pickle.load(f)
    vocab = synth.Vocab(0)
    for gi in range(2):
        sg = sgs[gi]
        ds = bd.BaseDataset()
        ds.vocab = {"pred_name_to_idx": dict(vocab.pred_ids), "pred_idx_to_name": list(vocab.pred_names)}
        ds.learned_converse, ds.learned_transitivity = True, True
        W = synth.make_conv_weights(vocab, seed=50 + gi)
        ds.converse_candidates_weights = W
        base = np.array(sg["relationships"]).astype(np.int64)
        np.random.seed(5 + gi)
        r_trip, r_counts, r_type = ds.add_learnt_triplets([list(x) for x in base], len(sg["objects"]))
        uniforms = np.random.RandomState(5 + gi).random_sample(len(base) + 8)
        o_trip, o_counts, o_type, used = ocanon.add_learnt_triplets(
            base, vocab.num_preds, vocab.meta_ids, W, True, True, uniforms)
        assert (o_trip == np.asarray(r_trip)).all() and (o_counts == r_counts).all()
        out["pkl%d_base" % gi] = base
        out["pkl%d_uniforms" % gi] = uniforms[:used]
        out["pkl%d_seed" % gi] = np.array(50 + gi)
        out["pkl%d_triplets" % gi] = np.asarray(r_trip).astype(np.int64)
        out["pkl%d_counts" % gi] = r_counts
        out["pkl%d_type" % gi] = np.asarray(r_type).astype(np.int64)
        print("pkl graph %d: O=%d base=%d -> T'=%d" % (gi, len(sg["objects"]), len(base), len(r_trip)))
    np.savez_compressed(os.path.join(OUT, "canon.npz"), **out)


# ----------------------------------------------------------------------------
def layer_inputs(seed=0, B=3, O=7, T=24, D=128, P=8):
    """Hand-shaped padded batch for one GraphTripleConv layer: padded rows, all four
    edge types, a self loop, a duplicate triple and an object no triple touches."""
    obj = synth.det_tensor((B, O, D), seed + 1, 1.0)
    pred = synth.det_tensor((B, T, D), seed + 2, 1.0)
    s = synth.det_int(B * T, seed + 3, 0, O - 2).reshape(B, T)      # object O-1 never used -> count 0
    o = synth.det_int(B * T, seed + 4, 0, O - 2).reshape(B, T)
    p = synth.det_int(B * T, seed + 5, 1, P - 1).reshape(B, T)
    ty = (synth.det_uniform(B * T, seed + 6).reshape(B, T) * 4).astype(np.int64)   # 0..3
    ty[:, :6] = 0
    s[0, 1], o[0, 1] = 2, 2                      # self loop
    s[0, 3], o[0, 3], p[0, 3], ty[0, 3] = s[0, 2], o[0, 2], p[0, 2], ty[0, 2]    # duplicate
    n_real = [T, T - 5, T - 11]
    for b in range(B):                            # collate padding: [0, __padding__, 0], type 0
        s[b, n_real[b]:], o[b, n_real[b]:], p[b, n_real[b]:], ty[b, n_real[b]:] = 0, 0, 0, 0
    return obj, pred, s, o, p, ty


def layer_state(seed, D=128, H=512, P=8):
    st = {}
    k = seed * 17

    def lin(name, of, inf):
        nonlocal k
        k += 2
        st[name + ".weight"] = synth.det_tensor((of, inf), k, float(np.sqrt(6.0 / inf)))
        st[name + ".bias"] = synth.det_tensor((of,), k + 1, float(1.0 / np.sqrt(inf)))
    lin("net1.0", H, 3 * D)
    lin("net1.2", 2 * H + D, H)
    lin("net2.0", H, H)
    lin("net2.2", D, H)
    st["predicates_transitive_weights"] = synth.det_tensor((P,), k + 9, 1.0)
    return st


GRAD_STRIDE = 97   # weight gradients are stored subsampled (flat[::GRAD_STRIDE]) to keep fixtures small


def gen_gconv(rgraph):
    D, H, P = 128, 512, 8
    obj, pred, s, o, p, ty = layer_inputs()
    st = layer_state(3)
    w_trans = torch.nn.Parameter(t(st["predicates_transitive_weights"]).clone())
    layer = rgraph.GraphTripleConv(D, D, D, D, H, 1, predicates_transitive_weights=w_trans)
    sd = {k: t(v) for k, v in st.items() if k != "predicates_transitive_weights"}
    layer.load_state_dict(sd, strict=False)
    obj_t = t(obj).clone().requires_grad_(True)
    pred_t = t(pred).clone().requires_grad_(True)
    edges = torch.stack([t(s), t(o)], dim=-1)
    ind = t(p) != 0
    new_obj, new_p = layer(obj_t, pred_t, edges, ind, t(ty), t(p))
    g_obj = t(synth.det_tensor(tuple(new_obj.shape), 41, 1.0))
    g_p = t(synth.det_tensor(tuple(new_p.shape), 42, 1.0))
    loss = (new_obj * g_obj).sum() + (new_p * g_p).sum()
    loss.backward()
    out = dict(new_obj=new_obj.detach().numpy(), new_p=new_p.detach().numpy(),
               d_obj=obj_t.grad.numpy(), d_pred=pred_t.grad.numpy(), d_w_trans=w_trans.grad.numpy())
    for name, prm in layer.named_parameters():
        if name == "predicates_transitive_weights":
            continue
        gflat = prm.grad.numpy().reshape(-1)
        out["dsub_" + name] = gflat[::GRAD_STRIDE].copy()
        out["dnorm_" + name] = np.array(np.linalg.norm(gflat.astype(np.float64)))
    # oracle restatement reproduces the reference (same ops, same order -> tight tolerance)
    st_t = {k: t(v).clone().requires_grad_(True) for k, v in st.items()}
    oo = t(obj).clone().requires_grad_(True)
    pp = t(pred).clone().requires_grad_(True)
    o_obj, o_p = ograph.graph_triple_conv(st_t, "", oo, pp, edges, ind, t(ty), t(p),
                                          st_t["predicates_transitive_weights"], H, D)
    ((o_obj * g_obj).sum() + (o_p * g_p).sum()).backward()
    assert torch.allclose(o_obj, new_obj, rtol=1e-6, atol=1e-6), "oracle GraphTripleConv fwd"
    assert torch.allclose(o_p, new_p, rtol=1e-6, atol=1e-6)
    assert torch.allclose(oo.grad, obj_t.grad, rtol=1e-5, atol=1e-6), "oracle GraphTripleConv bwd"
    assert torch.allclose(st_t["predicates_transitive_weights"].grad, w_trans.grad, rtol=1e-5, atol=1e-6)
    np.savez_compressed(os.path.join(OUT, "gconv_layer.npz"), **out)
    print("gconv layer: new_obj", tuple(new_obj.shape), "new_p", tuple(new_p.shape))


def model_batch(vocab, graphs, W_conv, seed, conv=True, trans=True):
    """Canonicalise each synthetic graph with the oracle, then pad like the reference collate
    (packed_coco.py:385-478)."""
    per = []
    for gi, g in enumerate(graphs):
        uni = synth.det_uniform(len(g.triplets) * 2 + 8, seed * 1000 + gi)
        trip, _, ty, _ = ocanon.add_learnt_triplets(g.triplets, vocab.num_preds, vocab.meta_ids, W_conv,
                                                    conv, trans, uni)
        per.append((g, trip, ty))
    Omax = max(len(g.objs) for g, _, _ in per)
    Tmax = max(len(tr) for _, tr, _ in per)
    B, A = len(per), vocab.num_attributes
    objs = np.zeros((B, Omax, A), np.int64)
    boxes = -np.ones((B, Omax, 4), np.float32)
    trips = np.zeros((B, Tmax, 3), np.int64)
    trips[:, :, 1] = vocab.padding_id
    types = np.zeros((B, Tmax), np.int64)
    for b, (g, tr, ty) in enumerate(per):
        objs[b, :len(g.objs)] = g.objs
        boxes[b, :len(g.boxes)] = g.boxes
        trips[b, :len(tr)] = tr
        types[b, :len(ty)] = ty
    return objs, boxes, trips, types


def gen_model(rmodel):
    vocab = synth.Vocab(0)
    graphs = synth.make_graphs(4, 11, 3, 8, vocab, include_dummies=True)
    W = synth.make_conv_weights(vocab, 1)
    objs, boxes, trips, types = model_batch(vocab, graphs, W, seed=3)
    st = synth.make_state(vocab, seed=2)
    opt = argparse.Namespace(
        vocab={"attributes": {"objects": {str(i): i for i in range(vocab.num_obj_classes)}},
               "pred_idx_to_name": vocab.pred_names, "pred_name_to_idx": vocab.pred_ids},
        image_size=(64, 64), layout_noise_dim=0, mask_noise_dim=0, embedding_dim=128, gconv_dim=128,
        gconv_hidden_dim=512, gconv_pooling="avg", gconv_num_layers=5, mlp_normalization="none",
        mask_size=0, learned_init="uniform")
    model = rmodel.Sg2LayoutModel(opt)
    missing = model.load_state_dict({k: t(v) for k, v in st.items()}, strict=False)
    # every GraphTripleConv registers the shared transitive weights a second time (graph.py:42)
    assert all("predicates_transitive_weights" in k for k in missing.missing_keys), missing
    obj_vecs, boxes_pred, masks_pred = model(t(objs), t(trips), t(types))
    assert masks_pred is None
    real = t((boxes >= 0).all(-1))
    loss = boxes_pred.pow(2).sum() + (obj_vecs * t(synth.det_tensor(tuple(obj_vecs.shape), 5, 1.0))).sum()
    loss.backward()
    out = dict(objs=objs, boxes=boxes, triplets=trips, types=types,
               obj_vecs=obj_vecs.detach().numpy(), boxes_pred=boxes_pred.detach().numpy(),
               loss=np.array(loss.item()))
    for name, prm in model.named_parameters():
        if "predicates_transitive_weights" in name or prm.grad is None:
            continue
        gflat = prm.grad.numpy().reshape(-1)
        if gflat.size <= 4096:
            out["d_" + name] = prm.grad.numpy().copy()
        else:
            out["dsub_" + name] = gflat[::GRAD_STRIDE].copy()
        out["dnorm_" + name] = np.array(np.linalg.norm(gflat.astype(np.float64)))
    assert model.converse_candidates_weights.grad is None
    # oracle check
    st_t = {k: t(v).clone().requires_grad_(True) for k, v in st.items()}
    o_vecs, o_boxes = ograph.sg2layout_forward(st_t, t(objs), t(trips), t(types), vocab.padding_id)
    assert torch.allclose(o_vecs, obj_vecs, rtol=1e-5, atol=1e-6), "oracle Sg2LayoutModel"
    assert torch.allclose(o_boxes, boxes_pred, rtol=1e-5, atol=1e-6)
    np.savez_compressed(os.path.join(OUT, "sg2layout_model.npz"), **out)
    print("model: B=%d O=%d T=%d loss=%.6f" % (objs.shape[0], objs.shape[1], trips.shape[1], loss.item()))
    _ = real


def gen_layout(rlayout, rbil):
    import torch.nn.functional as F
    out = {}
    demo_vecs = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    demo_boxes = np.array([[0.25, 0.125, 0.5, 0.875], [0, 0, 1, 0.25], [0.6125, 0, 0.875, 1],
                           [0, 0.8, 1, 1.0], [0.25, 0.125, 0.5, 0.875], [0.6125, 0, 0.875, 1]], np.float32)
    plus = np.array([[0, 0, 1, 0, 0], [0, 1, 1, 1, 0], [1, 1, 1, 1, 1], [0, 1, 1, 1, 0], [0, 0, 1, 0, 0]], np.float32)
    ring = np.array([[0, 0, 1, 0, 0], [0, 1, 0, 1, 0], [1, 0, 0, 0, 1], [0, 1, 0, 1, 0], [0, 0, 1, 0, 0]], np.float32)
    demo_masks = np.stack([plus, ring, plus, plus, plus, plus])              # layout.py:215-258
    out.update(demo_vecs=demo_vecs, demo_boxes=demo_boxes, demo_masks=demo_masks)

    def run(tag, vecs, boxes, masks, H, W, test_mode=False, float_masks=False, legacy=False):
        v = t(vecs).clone().requires_grad_(True)
        b = t(boxes).clone().requires_grad_(True)
        m = None
        orig = F.grid_sample
        if legacy:   # torch<=1.2 semantics the checkpoints were trained under
            F.grid_sample = lambda *a, **k: orig(*a, **{**k, "align_corners": True})
        try:
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                if masks is None:
                    y = rlayout.boxes_to_layout(v, b, H, W)
                    oy = olayout.boxes_to_layout(t(vecs), t(boxes), H, W, align_corners=legacy)
                else:
                    m = t(masks).clone()
                    if float_masks:
                        m.requires_grad_(True)
                    y = rlayout.masks_to_layout(v, b, m, H, W, test_mode=test_mode)
                    oy = olayout.masks_to_layout(t(vecs), t(boxes), t(masks), H, W, test_mode=test_mode,
                                                 align_corners=legacy)
        finally:
            F.grid_sample = orig
        assert torch.equal(oy, y.detach()), "oracle layout %s" % tag
        out[tag + "_out"] = y.detach().numpy()
        if not test_mode:
            g = t(synth.det_tensor(tuple(y.shape), 77, 1.0))
            (y * g).sum().backward()
            out[tag + "_dvecs"] = v.grad.numpy()
            out[tag + "_dboxes"] = b.grad.numpy()
            if float_masks:
                out[tag + "_dmasks"] = m.grad.numpy()

    run("demo_boxes64", demo_vecs, demo_boxes, None, 64, 64)
    run("demo_masks64", demo_vecs, demo_boxes, demo_masks, 64, 64, float_masks=True)
    run("demo_masks64_test", demo_vecs, demo_boxes, demo_masks, 64, 64, test_mode=True)
    run("demo_boxes64_legacy", demo_vecs, demo_boxes, None, 64, 64, legacy=True)
    run("demo_masks64_legacy", demo_vecs, demo_boxes, demo_masks, 64, 64, legacy=True, float_masks=True)
    # random objects: D=16, rectangular canvas, boxes partly outside the canvas, int64 masks
    vocab = synth.Vocab(0)
    g = synth.make_graph(123, 7, 7, vocab, include_dummies=False, mask_size=16)
    vecs = synth.det_tensor((7, 16), 9, 1.0)
    boxes = g.boxes.copy()
    boxes[0] = [-0.2, 0.1, 0.5, 0.4]
    boxes[1] = [0.7, 0.8, 0.6, 0.5]
    boxes[2] = [1.5, 1.5, 0.2, 0.2]          # entirely outside: contributes zeros
    boxes[3] = [0.3, 0.3, 0.02, 0.015]       # sub-pixel box
    out.update(rnd_vecs=vecs, rnd_boxes=boxes, rnd_masks=g.masks)
    run("rnd_boxes", vecs, boxes, None, 32, 48)
    run("rnd_masks", vecs, boxes, g.masks, 32, 48)
    run("rnd_masks_f", vecs, boxes, synth.det_uniform(7 * 16 * 16, 5).reshape(7, 16, 16).astype(np.float32),
        40, 40, float_masks=True)
    out["rnd_masks_f_in"] = synth.det_uniform(7 * 16 * 16, 5).reshape(7, 16, 16).astype(np.float32)
    run("rnd_masks_test", vecs, boxes, g.masks, 32, 48, test_mode=True)
    # single object / avg pooling value
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import io
        import contextlib
        with contextlib.redirect_stdout(io.StringIO()):
            out["rnd_boxes_avg_out"] = rlayout.boxes_to_layout(t(vecs), t(boxes), 32, 48, pooling="avg").numpy()

    # crop_bbox (bilinear.py:65-94)
    imgs = synth.det_tensor((3, 3, 24, 20), 31, 1.0)
    bb = np.array([[0.1, 0.2, 0.5, 0.6], [0.0, 0.0, 1.0, 1.0], [0.6, 0.5, 0.7, 0.8]], np.float32)
    im = t(imgs).clone().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        crops = rbil.crop_bbox(im, t(bb), 8, 12)
    gc = t(synth.det_tensor(tuple(crops.shape), 32, 1.0))
    (crops * gc).sum().backward()
    assert torch.equal(olayout.crop_bbox(t(imgs), t(bb), 8, 12), crops.detach())
    out.update(crop_imgs=imgs, crop_boxes=bb, crop_out=crops.detach().numpy(), crop_dimgs=im.grad.numpy())
    np.savez_compressed(os.path.join(OUT, "layout.npz"), **out)
    print("layout fixtures:", len(out), "arrays")


def gen_collate():
    """packed_coco.py:385-478 / packed_vg.py:147-229 on seeded samples (the dataset modules import h5py / PIL /
    pycocotools at module level: absent here, stubbed, never called by the collate functions)."""
    for name in ("h5py", "PIL", "PIL.Image", "torchvision", "torchvision.transforms", "pycocotools", "pycocotools.mask",
                 "skimage", "skimage.transform", "imageio", "cv2"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    import sg2im.data.packed_coco as pc
    import sg2im.data.packed_vg as pv
    from oracle import collate as ocollate
    out = {}
    for ci, (fn, with_masks, na, seed) in enumerate([(pc.coco_collate_fn, True, 1, 3), (pv.vg_collate_fn, False, 1, 4),
                                                      (pc.coco_collate_fn, False, 3, 5)]):
        vocab = synth.Vocab(0, num_attributes=na)
        rv = {"pred_name_to_idx": vocab.pred_ids}
        samples = gi.collate_samples(vocab, seed, 5, with_masks, 6)
        ref = fn(rv, samples)
        mine = ocollate.padded_collate(rv, samples, with_masks=fn is pc.coco_collate_fn)
        for k, (a, b) in enumerate(zip(ref, mine)):
            assert (a is None and b is None) or (a.dtype == b.dtype and torch.equal(a, b)), ("oracle collate", ci, k)
        out["c%d_spec" % ci] = np.array([int(with_masks), na, seed, int(fn is pc.coco_collate_fn)])
        for k, a in enumerate(ref):
            if a is not None:
                out["c%d_out%d" % (ci, k)] = a.numpy()
    out["num_cases"] = np.array(3)
    np.savez_compressed(os.path.join(OUT, "collate.npz"), **out)


def reference_canonicalize(bd, vocab, g, W, seed, conv=True, trans=True, dummies=True):
    """The unmodified ``BaseDataset.add_learnt_triplets`` on one synthetic graph, with numpy's global RNG seeded the way
    the oracle's explicit draws are (RandomState(seed).random_sample)."""
    ds = bd.BaseDataset()
    ds.vocab = {"pred_name_to_idx": dict(vocab.pred_ids), "pred_idx_to_name": list(vocab.pred_names),
                "object_name_to_idx": {"__image__": vocab.image_obj_id},
                "attributes": {"objects": {"__image__": vocab.image_obj_id}}}
    ds.include_dummies = dummies
    ds.learned_converse, ds.learned_transitivity = bool(conv), bool(trans)
    ds.converse_candidates_weights = W
    np.random.seed(seed)
    r_trip, r_counts, r_type = ds.add_learnt_triplets([list(x) for x in g.triplets], len(g.objs))
    uniforms = np.random.RandomState(seed).random_sample(len(g.triplets) * 2 + 8)
    o_trip, o_counts, o_type, used = ocanon.add_learnt_triplets(g.triplets, vocab.num_preds, vocab.meta_ids, W,
                                                                bool(conv), bool(trans), uniforms)
    assert (o_trip == np.asarray(r_trip).astype(np.int64)).all() and (o_type == np.asarray(r_type)).all()
    return np.asarray(r_trip).astype(np.int64), np.asarray(r_type).astype(np.int64), uniforms[:max(used, 1)]


def reference_bbox_loss(boxes_pred, boxes, objs, weight=10.0):
    """``Pix2PixModel.compute_generator_loss`` (pix2pix_model.py:65-143) of the unmodified reference with generation
    switched off: only the box term runs."""
    import sg2im.pix2pix_model as pm
    me = types.SimpleNamespace(opt=argparse.Namespace(skip_graph_model=False, skip_generation=True,
                                                      bbox_pred_loss_weight=weight))
    batch = (None, objs, boxes, None, None, None, None, None)
    out = pm.Pix2PixModel.compute_generator_loss(me, batch, (None, boxes_pred, None))
    return out["bbox_pred"], out["bbox_pred_all"]


def gen_box_loss():
    """Box term of the generator loss on padded batches: single-attribute (VG / COCO) and 4-attribute (CLEVR) objects,
    errors on both sides of the smooth-L1 knee."""
    from oracle import step as ostep
    out = {}
    for ci, (A, B, O, seed) in enumerate([(1, 5, 9, 1), (4, 3, 12, 2), (1, 64, 31, 3)]):
        objs = synth.det_int(B * O * A, seed * 11 + 1, 0 if A > 1 else 1, 7).reshape(B, O, A)
        n_real = synth.det_int(B, seed * 11 + 2, 1, O - 1)
        boxes = synth.det_tensor((B, O, 4), seed * 11 + 3, 0.5) + np.float32(0.5)
        for b in range(B):
            objs[b, n_real[b]:] = 0                      # __image__ dummy + collate padding are all-zero rows
            boxes[b, n_real[b]:] = -1.0
            if A > 1:
                objs[b, :n_real[b], 0] = np.maximum(objs[b, :n_real[b], 0], 1)
        pred = boxes + synth.det_tensor((B, O, 4), seed * 11 + 4, 2.0)
        bp = t(pred).clone().requires_grad_(True)
        loss, loss_all = reference_bbox_loss(bp, t(boxes), t(objs))
        (3.0 * loss).backward()
        o_loss, o_all = ostep.bbox_pred_loss(t(pred), t(boxes), t(objs))
        assert torch.allclose(o_loss, loss.detach(), rtol=1e-6, atol=0) and torch.allclose(o_all, loss_all.detach(), rtol=1e-6, atol=0)
        out["c%d_spec" % ci] = np.array([A, B, O, seed])
        out["c%d_loss" % ci] = np.array(loss.item(), np.float64)
        out["c%d_loss_all" % ci] = loss_all.detach().numpy()
        out["c%d_dpred" % ci] = bp.grad.numpy()
    out["num_cases"] = np.array(3)
    np.savez_compressed(os.path.join(OUT, "box_loss.npz"), **out)
    print("box loss fixtures:", len(out), "arrays")


def _store_model_run(out, model, obj_vecs, boxes_pred, loss):
    out.update(obj_vecs=obj_vecs.detach().numpy(), boxes_pred=boxes_pred.detach().numpy(), loss=np.array(loss.item()))
    for name, prm in model.named_parameters():
        if "predicates_transitive_weights" in name or prm.grad is None:
            continue
        gflat = prm.grad.numpy().reshape(-1)
        if gflat.size <= 4096:
            out["d_" + name] = prm.grad.numpy().copy()
        else:
            out["dsub_" + name] = gflat[::GRAD_STRIDE].copy()
        out["dnorm_" + name] = np.array(np.linalg.norm(gflat.astype(np.float64)))


def _gen_shape_model(bd, rmodel, case, fname):
    vocab, graphs, W, seeds, st, opt = case
    canon = []
    for g, seed in zip(graphs, seeds):
        tr, ty, _ = reference_canonicalize(bd, vocab, g, W, seed)
        canon.append((tr, ty))
    objs, boxes, trips, types_ = bc.pad_batch(vocab, graphs, canon)
    model = rmodel.Sg2LayoutModel(opt)
    missing = model.load_state_dict({k: t(v) for k, v in st.items()}, strict=False)
    assert all("predicates_transitive_weights" in k for k in missing.missing_keys), missing
    obj_vecs, boxes_pred, _ = model(t(objs), t(trips), t(types_))
    bl, _ = reference_bbox_loss(boxes_pred, t(boxes), t(objs))
    loss = bl + (obj_vecs * t(bc.obj_grad(obj_vecs.shape))).sum() * 1e-2
    loss.backward()
    out = dict(n_canon=np.array([len(tr) for tr, _ in canon]), types_sum=np.array(int(types_.sum())),
               trip_hash=np.array(int((trips.astype(np.int64) * np.arange(1, trips.size + 1).reshape(trips.shape)).sum()
                                      % (2 ** 61 - 1))))
    _store_model_run(out, model, obj_vecs, boxes_pred, loss)
    st_t = {k: t(v).clone() for k, v in st.items()}
    o_vecs, o_boxes = ograph.sg2layout_forward(st_t, t(objs), t(trips), t(types_), vocab.padding_id)
    assert torch.allclose(o_vecs, obj_vecs, rtol=1e-5, atol=1e-6) and torch.allclose(o_boxes, boxes_pred, rtol=1e-5, atol=1e-6)
    np.savez_compressed(os.path.join(OUT, fname), **out)
    print("%s: B=%d O=%d T=%d (real triples %s) loss=%.6f" % (fname, objs.shape[0], objs.shape[1], trips.shape[1],
                                                               [len(tr) for tr, _ in canon], loss.item()))


def gen_cfg2_model(bd, rmodel):
    """BASELINE config 2 shapes through the UNMODIFIED reference: 8 VG-like graphs (3-30 objects, P = 50), canonicalized
    by the reference's add_learnt_triplets (learned converse + transitive), padded by the reference's collate rule,
    Sg2LayoutModel forward, box loss of compute_generator_loss + a seeded linear functional of obj_vecs, backward.
    Inputs: tests/baseline_cases.py."""
    _gen_shape_model(bd, rmodel, bc.cfg2_case(), "cfg2_model.npz")


def gen_cfg4_model(bd, rmodel):
    """BASELINE config 4 shapes through the UNMODIFIED reference: 2 CLEVR-like graphs of 32-64 objects with 4
    attributes x embedding 32 (attribute_fc_gen in front of the GCN), closure-dense canonicalized triples."""
    _gen_shape_model(bd, rmodel, bc.cfg4_case(), "cfg4_model.npz")


def gen_cfg3_layout(rlayout):
    """BASELINE config 3 / 4 canvas shapes through the UNMODIFIED reference: one 256 x 256 image, D = 128, 16 x 16 masks
    (train path with gradient wrt vecs; test_mode occlusion path).  The 32 MiB canvases are stored subsampled
    (every 8th row and column, which are also rows / columns of the SPADE pyramid levels)."""
    vocab = synth.Vocab(0)
    out = {}
    g = synth.make_graph(777, 8, 8, vocab, include_dummies=False, mask_size=16)
    vecs = synth.det_tensor((8, 128), 12, 1.0)
    gsub = synth.det_tensor((1, 128, 32, 32), 13, 1.0)
    v = t(vecs).clone().requires_grad_(True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        y = rlayout.masks_to_layout(v, t(g.boxes), t(g.masks), 256, 256)
        yt = rlayout.masks_to_layout(t(vecs), t(g.boxes), t(g.masks), 256, 256, test_mode=True)
        oy = olayout.masks_to_layout(t(vecs), t(g.boxes), t(g.masks), 256, 256)
        oyt = olayout.masks_to_layout(t(vecs), t(g.boxes), t(g.masks), 256, 256, test_mode=True)
    assert torch.equal(oy, y.detach()) and torch.equal(oyt, yt)
    # upstream gradient: non-zero on the 8-strided lattice only, so the fixture stays small and regenerable
    G = torch.zeros_like(y)
    G[:, :, ::8, ::8] = t(gsub)
    (y * G).sum().backward()
    out.update(train_sub=y.detach()[:, :, ::8, ::8].numpy(), test_sub=yt[:, :, ::8, ::8].numpy(),
               train_sum=np.array(y.detach().double().sum().item()), test_sum=np.array(yt.double().sum().item()),
               train_abs=np.array(y.detach().double().abs().sum().item()), dvecs=v.grad.numpy())
    np.savez_compressed(os.path.join(OUT, "cfg3_layout.npz"), **out)
    print("cfg3 layout: canvas", tuple(y.shape))


def gen_pyramid(rlayout):
    """SPADE-side consumers of the canvas (spade/models/networks/generator.py:99, normalization.py:102): the nearest
    resizes every SPADE block applies to the full-resolution canvas, for a 64 x 64 canvas of 3 images."""
    import torch.nn.functional as F
    vocab = synth.Vocab(0)
    graphs = synth.make_graphs(3, 31, 2, 6, vocab, include_dummies=False)
    canv = []
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i, g in enumerate(graphs):
            canv.append(rlayout.boxes_to_layout(t(synth.det_tensor((len(g.boxes), 16), 50 + i, 1.0)), t(g.boxes), 64, 64))
    seg = torch.cat(canv, 0)
    out = {"seg": seg.numpy()}
    out["head"] = F.interpolate(seg, size=(2, 2)).numpy()                              # generator.py:99 (sh, sw)
    for s_ in (4, 8, 16, 32):
        out["l%d" % s_] = F.interpolate(seg, size=(s_, s_), mode="nearest").numpy()    # normalization.py:102
    np.savez_compressed(os.path.join(OUT, "pyramid.npz"), **out)


def gen_crop_jj(rbil):
    """crop_bbox(backend='jj') = bilinear_sample (bilinear.py:97-152) of the unmodified reference, fwd + d/dfeats;
    boxes inside the image, touching its border and reaching beyond it (clamped taps)."""
    imgs = synth.det_tensor((4, 3, 24, 20), 31, 1.0)
    bb = np.array([[0.1, 0.2, 0.5, 0.6], [0.0, 0.0, 1.0, 1.0], [0.6, 0.5, 0.7, 0.8], [0.93, 0.9, 0.05, 0.08]], np.float32)
    im = t(imgs).clone().requires_grad_(True)
    crops = rbil.crop_bbox(im, t(bb), 8, 12, backend="jj")
    gc = t(synth.det_tensor(tuple(crops.shape), 32, 1.0))
    (crops * gc).sum().backward()
    oc = olayout.crop_bbox_jj(t(imgs), t(bb), 8, 12)
    assert torch.allclose(oc, crops.detach(), rtol=0, atol=1e-6), "oracle crop_bbox jj"
    np.savez_compressed(os.path.join(OUT, "crop_jj.npz"), imgs=imgs, boxes=bb, out=crops.detach().numpy(),
                        dimgs=im.grad.numpy())
    print("crop jj:", tuple(crops.shape))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated fixture names (default: all)")
    only = [x for x in ap.parse_args().only.split(",") if x]
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)
    gu, bd, rgraph, rmodel, rlayout, rbil = import_reference()
    jobs = [("canon", lambda: gen_canon(gu, bd)), ("gconv_layer", lambda: gen_gconv(rgraph)),
            ("sg2layout_model", lambda: gen_model(rmodel)), ("layout", lambda: gen_layout(rlayout, rbil)),
            ("collate", gen_collate), ("box_loss", gen_box_loss), ("cfg2_model", lambda: gen_cfg2_model(bd, rmodel)),
            ("cfg4_model", lambda: gen_cfg4_model(bd, rmodel)), ("cfg3_layout", lambda: gen_cfg3_layout(rlayout)),
            ("pyramid", lambda: gen_pyramid(rlayout)), ("crop_jj", lambda: gen_crop_jj(rbil))]
    for name, fn in jobs:
        if not only or name in only:
            fn()
    for f in sorted(os.listdir(OUT)):
        print("%-28s %8.1f KB" % (f, os.path.getsize(os.path.join(OUT, f)) / 1024))


if __name__ == "__main__":
    main()
