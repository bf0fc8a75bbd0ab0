"""CUDA canonicalization (csrc/canon.cu) vs the reference's golden vectors: bit-exact edge sets,
edge types and converse counts."""
import numpy as np
import pytest
import torch

from canonicalsg2im_b200 import synth
from tests.util import t

pytestmark = pytest.mark.gpu


def _case_graph(spec):
    nb, na, n0, n1, dummies, clevr, conv, trans, gseed, ci, used = [int(x) for x in spec]
    vocab = synth.Vocab(nb, num_attributes=na)
    g = synth.make_graph(gseed, n0, n1, vocab, include_dummies=bool(dummies), box_mode="clevr" if clevr else "coco")
    return vocab, g, bool(conv), bool(trans), ci


def test_closure_and_reduction_golden(golden):
    from canonicalsg2im_b200 import canonicalize as C
    g = golden("canon")
    for i in range(int(g["num_adj"])):
        a = t(g["adj%d" % i])
        assert (C.closure(a).cpu().numpy().astype(bool) == g["adj%d_path" % i]).all()
        assert (C.closure(a, reduce=True).cpu().numpy().astype(bool) == g["adj%d_min" % i]).all()
    kat = t(g["kat_adj"].astype(np.uint8))                      # graphs_utils.py:159-174
    m = C.closure(kat, reduce=True).cpu().numpy()
    assert (np.argwhere(m) == [[0, 3], [1, 2], [3, 1]]).all()


def test_add_learnt_triplets_golden(golden):
    from canonicalsg2im_b200 import canonicalize as C
    gd = golden("canon")
    for c in range(int(gd["num_cases"])):
        vocab, g, conv, trans, ci = _case_graph(gd["c%d_spec" % c])
        W = synth.make_conv_weights(vocab, seed=ci)
        trip, counts, types = C.add_learnt_triplets(g.triplets, len(g.objs), vocab.num_preds, vocab.meta_ids, W,
                                                    conv, trans, gd["c%d_uniforms" % c])
        assert trip.shape == gd["c%d_triplets" % c].shape, "case %d" % c
        assert (trip == gd["c%d_triplets" % c]).all(), "case %d edges" % c
        assert (np.asarray(types) == gd["c%d_type" % c]).all(), "case %d types" % c
        assert (counts == gd["c%d_counts" % c]).all(), "case %d conv_counts" % c


def test_shipped_clevr_graphs(golden):
    from canonicalsg2im_b200 import canonicalize as C
    gd = golden("canon")
    vocab = synth.Vocab(0)
    for gi_ in range(2):
        W = synth.make_conv_weights(vocab, seed=int(gd["pkl%d_seed" % gi_]))
        base = gd["pkl%d_base" % gi_]
        O = int(base[:, [0, 2]].max()) + 1
        trip, counts, types = C.add_learnt_triplets(base, O, vocab.num_preds, vocab.meta_ids, W, True, True,
                                                    gd["pkl%d_uniforms" % gi_])
        assert (trip == gd["pkl%d_triplets" % gi_]).all()
        assert (counts == gd["pkl%d_counts" % gi_]).all()
        assert (np.asarray(types) == gd["pkl%d_type" % gi_]).all()


def test_batched_equals_oracle_cfg2_shape():
    """A cfg2-shaped batch (VG-like P=50, 3-30 objects, both learned flags) in ONE call vs the oracle per graph."""
    from canonicalsg2im_b200 import canonicalize as C
    from oracle import canon as ocanon
    vocab = synth.Vocab(42)
    graphs = synth.make_graphs(24, 5, 3, 30, vocab, include_dummies=True)
    W = synth.make_conv_weights(vocab, 9)
    tri_off = np.concatenate([[0], np.cumsum([len(g.triplets) for g in graphs])]).astype(np.int32)
    obj_off = np.concatenate([[0], np.cumsum([len(g.objs) for g in graphs])]).astype(np.int32)
    trip = np.concatenate([g.triplets for g in graphs])
    uni = synth.det_uniform(len(trip), 123)
    res = C.add_learnt_triplets_batched(t(trip), t(tri_off), t(obj_off), vocab.num_preds, vocab.meta_ids, W,
                                        True, True, t(uni))
    parts = res.split()
    for gidx, g in enumerate(graphs):
        o_t, o_c, o_y, _ = ocanon.add_learnt_triplets(g.triplets, vocab.num_preds, vocab.meta_ids, W, True, True,
                                                      uni[tri_off[gidx]:])
        r_t, r_c, r_y = parts[gidx]
        assert (r_t.cpu().numpy() == o_t).all(), "graph %d" % gidx
        assert (r_y.cpu().numpy() == o_y).all()
        assert (r_c.cpu().numpy() == o_c).all()


def test_properties_large():
    """cfg4-shaped graphs (64 objects): closure is idempotent and contains the input; minimal graph
    has the same closure — size-independent properties checked without the O(n^3) Python oracle."""
    from canonicalsg2im_b200 import canonicalize as C
    n, G = 64, 32
    a = (t(synth.det_uniform(G * n * n, 7).reshape(G, n, n)) < 0.03).to(torch.uint8)
    c = C.closure(a)
    assert (c >= a).all()
    assert torch.equal(C.closure(c), c)
    dag = torch.triu(a, diagonal=1)                               # acyclic: reduction is unique
    m = C.closure(dag, reduce=True)
    assert (m <= C.closure(dag)).all()
    assert torch.equal(C.closure(m), C.closure(dag))
    reach = (torch.linalg.matrix_power((dag.float() + torch.eye(n, device="cuda")), n) > 0)
    assert torch.equal(C.closure(dag).bool() | torch.eye(n, device="cuda").bool(), reach)


# ---------------------------------------------------------------------------- location triplets (base_dataset.py:35-87)
def _centers(g):
    n_real = len(g.centers)
    return np.concatenate([g.centers, np.zeros((len(g.boxes) - n_real, 2), np.float32)]).astype(np.float32)


def _gpu_location(graphs, vocab, max_objs=None, dummies=False):
    from canonicalsg2im_b200 import canonicalize as C
    boxes = np.concatenate([g.boxes for g in graphs]).astype(np.float32)
    cen = np.concatenate([_centers(g) for g in graphs])
    objs = np.concatenate([g.objs for g in graphs]).astype(np.int64)
    off = np.concatenate([[0], np.cumsum([len(g.boxes) for g in graphs])]).astype(np.int32)
    trip, tri_off = C.add_location_triplets_batched(t(boxes), t(cen), t(objs), t(off), vocab.image_obj_id, vocab.pred_ids,
                                                    max_objs_per_graph=max_objs,
                                                    in_image_pred=vocab.in_image_id if dummies else None)
    trip, tri_off = trip.cpu().numpy(), tri_off.cpu().numpy()
    return [trip[tri_off[i]:tri_off[i + 1]] for i in range(len(graphs))]


def test_location_triplets_golden(golden):
    """The reference's own output (location + dummy triplets of the golden graphs, oracle/make_golden.py) bit for bit."""
    from oracle import canon as ocanon
    gd = golden("canon")
    for c in range(int(gd["num_cases"])):
        nb, na, n0, n1, dummies, clevr = [int(x) for x in gd["c%d_spec" % c][:6]]
        vocab = synth.Vocab(nb, num_attributes=na)
        g = synth.make_graph(int(gd["c%d_spec" % c][8]), n0, n1, vocab, include_dummies=bool(dummies),
                             box_mode="clevr" if clevr else "coco")
        loc = _gpu_location([g], vocab)[0]
        dummy = ocanon.add_dummy_triplets(g.objs[:, 0], vocab.image_obj_id, vocab.in_image_id,
                                          include_dummies=len(g.boxes) > len(g.centers))
        got = np.concatenate([loc, np.array(dummy, dtype=np.int64).reshape(-1, 3)])
        assert (got == gd["c%d_base" % c]).all()
        # add_dummy_triplets on the device as well (csg_dummy_triplets_*): the whole reference list in one call
        both = _gpu_location([g], vocab, dummies=True)[0]
        assert both.shape == gd["c%d_base" % c].shape and (both == gd["c%d_base" % c]).all()


def test_location_and_dummy_triplets_batched():
    """Ragged batch, graphs with and without the __image__ object: location triplets followed by the
    [i, __in_image__, image] rows (base_dataset.py:141-150), graph by graph, against the oracle."""
    from oracle import canon as ocanon
    vocab = synth.Vocab(0)
    graphs = [synth.make_graph(9100 + i, 1, 40, vocab, include_dummies=(i % 4 != 1), box_mode="clevr" if i % 2 else "coco")
              for i in range(40)]
    got = _gpu_location(graphs, vocab, dummies=True)
    for g, trip in zip(graphs, got):
        has_img = len(g.boxes) > len(g.centers)
        ref = ocanon.add_location_triplets(g.boxes, _centers(g), g.objs[:, 0], vocab.image_obj_id, vocab.pred_ids)
        ref = list(ref) + list(ocanon.add_dummy_triplets(g.objs[:, 0], vocab.image_obj_id, vocab.in_image_id, has_img))
        ref = np.array(ref, dtype=np.int64).reshape(-1, 3)
        assert trip.shape == ref.shape and (trip == ref).all()


@pytest.mark.parametrize("n_min,n_max,mode", [(1, 4, "coco"), (3, 30, "coco"), (32, 64, "clevr"), (60, 100, "coco")])
def test_location_triplets_batched_equal_oracle(n_min, n_max, mode):
    """Ragged batches (single-object graphs, VG-sized graphs, CLEVR-sized graphs of 32-64 objects, > 64 objects = more
    than one bitset word per row) against the oracle port, graph by graph, bit-exact."""
    from oracle import canon as ocanon
    vocab = synth.Vocab(0)
    graphs = [synth.make_graph(5000 + 31 * n_max + i, n_min, n_max, vocab, include_dummies=(i % 3 != 0), box_mode=mode)
              for i in range(24)]
    got = _gpu_location(graphs, vocab)
    for g, trip in zip(graphs, got):
        ref = np.array(ocanon.add_location_triplets(g.boxes, _centers(g), g.objs[:, 0], vocab.image_obj_id, vocab.pred_ids),
                       dtype=np.int64).reshape(-1, 3)
        assert trip.shape == ref.shape and (trip == ref).all()


def test_location_triplets_oversize_graph_is_reported():
    from canonicalsg2im_b200 import _lib
    vocab = synth.Vocab(0)
    graphs = [synth.make_graph(77, 10, 12, vocab, include_dummies=True)]
    with pytest.raises(_lib.CsgError):
        _gpu_location(graphs, vocab, max_objs=4)
