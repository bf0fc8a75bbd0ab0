"""Oracle (oracle/canon.py) against the golden vectors produced by the reference
(scripts/graphs_utils.py, sg2im/data/base_dataset.py) — CPU only."""
import numpy as np
import pytest

from canonicalsg2im_b200 import synth
from oracle import canon


def test_reference_kat(golden):
    g = golden("canon")
    # graphs_utils.py:159-174
    assert (canon.triplets_to_adj(g["kat_triplets"]) == g["kat_adj"].astype(bool)).all()
    assert (canon.triplets_to_minimal(g["kat_triplets"]) == [[0, 1, 3], [1, 1, 2], [3, 1, 1]]).all()
    assert (canon.triplets_to_minimal(g["kat_triplets"]) == g["kat_minimal"]).all()
    assert (canon.current_and_transitive(g["kat_minimal"])[1] == g["kat_transitive"]).all()
    # a 2-cycle closes into self loops
    assert (canon.current_and_transitive([[0, 1, 1], [1, 1, 0]])[1] == g["cyc_transitive"]).all()
    assert (g["cyc_transitive"] == [[0, 1, 0], [1, 1, 1]]).all()


def test_fewer_than_three_is_identity():
    t = np.array([[0, 2, 1], [1, 2, 2]])
    assert (canon.triplets_to_minimal(t) == t).all()


def test_closure_and_reduction(golden):
    g = golden("canon")
    for i in range(int(g["num_adj"])):
        a = g["adj%d" % i]
        assert (canon.closure(a) == g["adj%d_path" % i]).all()
        assert (canon.minimal_graph(a) == g["adj%d_min" % i]).all()


def _case_graph(spec):
    nb, na, n0, n1, dummies, clevr, conv, trans, gseed, ci, used = [int(x) for x in spec]
    vocab = synth.Vocab(nb, num_attributes=na)
    g = synth.make_graph(gseed, n0, n1, vocab, include_dummies=bool(dummies),
                         box_mode="clevr" if clevr else "coco")
    return vocab, g, bool(conv), bool(trans), ci, used


def test_add_learnt_triplets_matches_reference(golden):
    gd = golden("canon")
    for c in range(int(gd["num_cases"])):
        vocab, g, conv, trans, ci, used = _case_graph(gd["c%d_spec" % c])
        W = synth.make_conv_weights(vocab, seed=ci)
        trip, counts, types, n_used = canon.add_learnt_triplets(
            g.triplets, vocab.num_preds, vocab.meta_ids, W, conv, trans, gd["c%d_uniforms" % c])
        assert n_used == used
        assert (trip == gd["c%d_triplets" % c]).all()
        assert (counts == gd["c%d_counts" % c]).all()
        assert (types == gd["c%d_type" % c]).all()


def test_location_and_dummy_triplets_match_reference(golden):
    gd = golden("canon")
    for c in range(int(gd["num_cases"])):
        vocab, g, _, _, _, _ = _case_graph(gd["c%d_spec" % c])
        n_real = len(g.centers)
        cen = np.concatenate([g.centers, np.zeros((len(g.boxes) - n_real, 2), np.float32)])
        loc = canon.add_location_triplets(g.boxes, cen, g.objs[:, 0], vocab.image_obj_id, vocab.pred_ids)
        dummy = canon.add_dummy_triplets(g.objs[:, 0], vocab.image_obj_id, vocab.in_image_id,
                                         include_dummies=len(g.boxes) > n_real)
        got = np.array(loc + dummy).reshape(-1, 3)
        assert (got == gd["c%d_base" % c]).all()


def test_shipped_clevr_graphs(golden):
    gd = golden("canon")
    vocab = synth.Vocab(0)
    for gi in range(2):
        W = synth.make_conv_weights(vocab, seed=int(gd["pkl%d_seed" % gi]))
        trip, counts, types, _ = canon.add_learnt_triplets(
            gd["pkl%d_base" % gi], vocab.num_preds, vocab.meta_ids, W, True, True, gd["pkl%d_uniforms" % gi])
        assert (trip == gd["pkl%d_triplets" % gi]).all()
        assert (counts == gd["pkl%d_counts" % gi]).all()
        assert (types == gd["pkl%d_type" % gi]).all()


@pytest.mark.parametrize("seed", range(6))
def test_synth_location_triplets_equal_oracle(seed):
    """The vectorised generator used for large synthetic batches equals the oracle port."""
    vocab = synth.Vocab(0)
    g = synth.make_graph(900 + seed, 3, 14, vocab, include_dummies=True)
    n_real = len(g.centers)
    cen = np.concatenate([g.centers, np.zeros((1, 2), np.float32)])
    loc = np.array(canon.add_location_triplets(g.boxes, cen, g.objs[:, 0], 0, vocab.pred_ids)).reshape(-1, 3)
    spatial = g.triplets[g.triplets[:, 1] >= 2]
    assert (spatial == loc).all()
    assert n_real + 1 == len(g.boxes)
