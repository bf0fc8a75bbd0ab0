"""Seeded inputs of the BASELINE-shaped parity cases (cfg2 / cfg4 models), shared by the GPU tests and
``oracle/make_golden.py`` (which ran the unmodified reference on exactly these tensors)."""
import argparse

import numpy as np

from canonicalsg2im_b200 import synth

CFG4_ATTR_SIZES = list(synth.CLEVR_ATTR_SIZES)


def pad_batch(vocab, graphs, canon):
    """packed_coco.py:385-478 padding of (objs, boxes, triplets, types)."""
    Omax = max(len(g.objs) for g in graphs)
    Tmax = max(len(tr) for tr, _ in canon)
    B, A = len(graphs), vocab.num_attributes
    objs = np.zeros((B, Omax, A), np.int64)
    boxes = -np.ones((B, Omax, 4), np.float32)
    trips = np.zeros((B, Tmax, 3), np.int64)
    trips[:, :, 1] = vocab.padding_id
    types = np.zeros((B, Tmax), np.int64)
    for b, (g, (tr, ty)) in enumerate(zip(graphs, canon)):
        objs[b, :len(g.objs)], boxes[b, :len(g.boxes)] = g.objs, g.boxes
        trips[b, :len(tr)], types[b, :len(ty)] = tr, ty
    return objs, boxes, trips, types


def model_opt(vocab, embedding_dim, attr_sizes):
    attrs = {"a%d" % i: {str(j): j for j in range(attr_sizes[i])} for i in range(vocab.num_attributes)}
    return argparse.Namespace(
        vocab={"attributes": attrs, "pred_idx_to_name": vocab.pred_names, "pred_name_to_idx": vocab.pred_ids},
        image_size=(64, 64), layout_noise_dim=0, mask_noise_dim=0, embedding_dim=embedding_dim, gconv_dim=128,
        gconv_hidden_dim=512, gconv_pooling="avg", gconv_num_layers=5, mlp_normalization="none",
        mask_size=0, learned_init="uniform")


def cfg2_case():
    """8 VG-like graphs (3-30 objects, P = 50): (vocab, graphs, converse weights, per-graph RNG seeds, state, opt)."""
    vocab = synth.Vocab(42)
    graphs = synth.make_graphs(8, 2024, 3, 30, vocab, include_dummies=True)
    W = synth.make_conv_weights(vocab, 7)
    seeds = [900 + i for i in range(len(graphs))]
    st = synth.make_state(vocab, seed=5)
    return vocab, graphs, W, seeds, st, model_opt(vocab, 128, [vocab.num_obj_classes])


def cfg4_case():
    """2 CLEVR-like graphs of 32-64 objects, 4 attributes x embedding 32 (attribute_fc_gen in front of the GCN)."""
    vocab = synth.Vocab(0, num_attributes=4)
    graphs = synth.make_graphs(2, 4040, 32, 64, vocab, include_dummies=True, box_mode="clevr")
    for g in graphs:
        for k, n in enumerate(CFG4_ATTR_SIZES):
            g.objs[:-1, k] = 1 + (g.objs[:-1, k] - 1) % (n - 1)
    W = synth.make_conv_weights(vocab, 9)
    seeds = [700 + i for i in range(len(graphs))]
    st = synth.make_state(vocab, embedding_dim=32, seed=8, attr_vocab_sizes=CFG4_ATTR_SIZES)
    return vocab, graphs, W, seeds, st, model_opt(vocab, 32, CFG4_ATTR_SIZES)


def canon_uniforms(g, seed):
    """The draws numpy's global RNG hands ``add_learnt_triplets`` after ``np.random.seed(seed)``."""
    return np.random.RandomState(seed).random_sample(len(g.triplets) * 2 + 8)


def obj_grad(shape):
    return synth.det_tensor(tuple(shape), 6, 1.0)
