"""Inputs of the golden fixtures, regenerated from integer seeds.

Mirrors ``oracle/make_golden.py`` (which ran the reference on exactly these
tensors); kept separate so that GPU tests can build the inputs without
importing the oracle."""
import numpy as np

from canonicalsg2im_b200 import synth

GRAD_STRIDE = 97


def layer_inputs(seed=0, B=3, O=7, T=24, D=128, P=8):
    obj = synth.det_tensor((B, O, D), seed + 1, 1.0)
    pred = synth.det_tensor((B, T, D), seed + 2, 1.0)
    s = synth.det_int(B * T, seed + 3, 0, O - 2).reshape(B, T)
    o = synth.det_int(B * T, seed + 4, 0, O - 2).reshape(B, T)
    p = synth.det_int(B * T, seed + 5, 1, P - 1).reshape(B, T)
    ty = (synth.det_uniform(B * T, seed + 6).reshape(B, T) * 4).astype(np.int64)
    ty[:, :6] = 0
    s[0, 1], o[0, 1] = 2, 2
    s[0, 3], o[0, 3], p[0, 3], ty[0, 3] = s[0, 2], o[0, 2], p[0, 2], ty[0, 2]
    n_real = [T, T - 5, T - 11]
    for b in range(B):
        s[b, n_real[b]:], o[b, n_real[b]:], p[b, n_real[b]:], ty[b, n_real[b]:] = 0, 0, 0, 0
    return obj, pred, s, o, p, ty


def layer_state(seed=3, D=128, H=512, P=8):
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


def layer_out_grads(shape_obj, shape_p):
    return synth.det_tensor(tuple(shape_obj), 41, 1.0), synth.det_tensor(tuple(shape_p), 42, 1.0)


def model_state():
    return synth.make_state(synth.Vocab(0), seed=2)


def model_obj_grad(shape):
    return synth.det_tensor(tuple(shape), 5, 1.0)


def layout_out_grad(shape):
    return synth.det_tensor(tuple(shape), 77, 1.0)


def crop_out_grad(shape):
    return synth.det_tensor(tuple(shape), 32, 1.0)


def collate_samples(vocab, seed, num, with_masks, P):
    """Per-sample tuples as the reference datasets' ``__getitem__`` returns them (packed_coco.py:373-383), seeded:
    ``(img, objs: dict attribute -> LongTensor[O], boxes, triplets, conv_counts, triplet_type, masks or None, image_id)``."""
    import torch
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x))
    samples = []
    graphs = synth.make_graphs(num, seed, 1, 9, vocab, include_dummies=True, mask_size=4 if with_masks else 0)
    for i, g in enumerate(graphs):
        O = len(g.boxes)
        img = t(synth.det_tensor((3, 4, 4), seed * 31 + i, 1.0))
        objs = {"a%d" % k: t(g.objs[:, k].astype(np.int64)) for k in range(g.objs.shape[1])}
        trip = t(g.triplets.astype(np.int64))
        ty = t((synth.det_uniform(len(g.triplets), seed * 7 + i) < 0.3).astype(np.int64))
        cc = t(synth.det_tensor((P, P + 1), seed * 13 + i, 1.0).astype(np.float64))
        masks = None
        if with_masks:
            m = np.zeros((O, 4, 4), np.int64)
            m[:len(g.masks)] = g.masks
            masks = t(m)
        samples.append((img, objs, t(g.boxes.astype(np.float32)), trip, cc, ty, masks, 1000 + i))
    return samples
