"""Ragged collate (canonicalsg2im_b200/collate.py) against the reference's padded collate
functions (packed_coco.py:385-478, packed_vg.py:147-229): golden outputs of the unmodified reference on seeded samples
(tests/golden/collate.npz), bit-exact.  Host logic only: runs without a GPU."""
import numpy as np
import torch

from canonicalsg2im_b200 import synth
from canonicalsg2im_b200.collate import ragged_collate_fn, ragged_to_padded, RaggedBatch
from oracle import collate as ocollate
from tests import golden_inputs as gi


def _cases(golden):
    gd = golden("collate")
    for c in range(int(gd["num_cases"])):
        with_masks, na, seed, is_coco = [int(x) for x in gd["c%d_spec" % c]]
        vocab = synth.Vocab(0, num_attributes=na)
        samples = gi.collate_samples(vocab, seed, 5, bool(with_masks), 6)
        ref = [gd["c%d_out%d" % (c, k)] if ("c%d_out%d" % (c, k)) in gd.files else None for k in range(8)]
        yield vocab, samples, bool(is_coco), ref


def _same(got, ref):
    for k, (a, b) in enumerate(zip(got, ref)):
        if b is None:
            assert a is None, k
        else:
            assert a is not None and tuple(a.shape) == b.shape and (a.numpy() == b).all(), k


def test_oracle_padded_collate_matches_reference(golden):
    for vocab, samples, is_coco, ref in _cases(golden):
        _same(ocollate.padded_collate({"pred_name_to_idx": vocab.pred_ids}, samples, with_masks=is_coco), ref)


def test_ragged_collate_pads_back_to_the_reference_batch(golden):
    for vocab, samples, is_coco, ref in _cases(golden):
        rb = ragged_collate_fn({"pred_name_to_idx": vocab.pred_ids}, samples)
        assert isinstance(rb, RaggedBatch) and rb["B"] == len(samples)
        off, toff = rb["obj_off"].tolist(), rb["tri_off"].tolist()
        assert off[0] == 0 and toff[0] == 0 and off[-1] == rb["objs"].shape[0] and toff[-1] == rb["triplets"].shape[0]
        assert rb["obj_off"].dtype == torch.int32 and rb["triplets"].dtype == torch.int64
        for b, s in enumerate(samples):                       # nothing is padded, every sample is a contiguous slice
            assert torch.equal(rb["boxes"][off[b]:off[b + 1]], s[2]) and torch.equal(rb["triplets"][toff[b]:toff[b + 1]], s[3])
        padded = list(ragged_to_padded(rb, vocab.pred_ids["__padding__"]))
        if not is_coco:
            padded[6] = None                                  # vg_collate_fn never returns masks (packed_vg.py:165)
        _same(padded, ref)
        assert rb["max_objs"] == ref[1].shape[1] and rb["max_triplets"] == ref[3].shape[1]


def test_ragged_batch_feeds_the_flat_model_inputs():
    """The flat batch is what forward_ragged / add_learnt_triplets_batched take: local ids stay below the graph size."""
    vocab = synth.Vocab(0)
    rb = ragged_collate_fn(None, gi.collate_samples(vocab, 9, 6, False, 6))
    off, toff = rb["obj_off"].tolist(), rb["tri_off"].tolist()
    for b in range(rb["B"]):
        t = rb["triplets"][toff[b]:toff[b + 1]]
        assert t.numel() == 0 or int(t[:, [0, 2]].max()) < off[b + 1] - off[b]
    moved = rb.to("cpu")
    assert moved is not rb and torch.equal(moved["objs"], rb["objs"]) and moved["max_objs"] == rb["max_objs"]
