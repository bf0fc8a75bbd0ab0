"""CPU model of the arithmetic of ``layout_bwd_colsum_kernel`` (csrc/layout.cu) against the
oracle's d/dvecs of ``boxes_to_layout`` (sg2im/layout.py:12-45 through autograd).

The kernel does not add ay(y)*ax(x)*dout(y, x) term by term: it keeps running column sums of the incoming gradient
over a row range and adds, per object, (ay(y) - ay(y + 1)) * <ax, C_y> on the rows where the row factor changes, the
last row of the range closing the sum.  This file restates exactly that in float32 numpy -- same order of operations
per (object, row range) -- and checks (a) that it visits far fewer (object, row) pairs than the direct sum and (b) that
its rounding stays inside the 1e-5 (relative to the tensor scale) contract for the range lengths the launcher picks
(16 / 32 / 64 rows), for sign-alternating and for same-sign gradients (the worst case for a running sum)."""
import numpy as np
import pytest
import torch

from canonicalsg2im_b200 import synth
from oracle import layout as olayout


def _separable_factors(boxes, H, W):
    """ay [O, H], ax [O, W] with S_o(y, x) = ay[o, y] * ax[o, x], read off the oracle itself: the canvas of one-hot
    vectors is S_o, and a box weight is an outer product whose factors reach 1 inside the box."""
    O = len(boxes)
    S = olayout.boxes_to_layout(torch.eye(O), torch.from_numpy(boxes), H, W)[0].numpy()       # [O, H, W]
    ay, ax = S.max(axis=2), S.max(axis=1)
    peak = S.reshape(O, -1).max(axis=1)
    ok = peak > 0
    # factors are defined up to a scale: ay * ax / peak reproduces S wherever the box has an interior pixel
    ax[ok] = ax[ok] / peak[ok, None]
    return S, ay.astype(np.float32), ax.astype(np.float32)


def _colsum_dvecs(dout, ay, ax, rows_per_range):
    """float32 restatement of the kernel: per row range, running column sums C and visits where ay changes."""
    D, H, W = dout.shape
    O = ay.shape[0]
    dv = np.zeros((O, D), np.float32)
    visits = 0
    for r0 in range(0, H, rows_per_range):
        r1 = min(H, r0 + rows_per_range)
        C = np.zeros((D, W), np.float32)
        part = np.zeros((O, D), np.float32)
        for y in range(r0, r1):
            C = (C + dout[:, y, :]).astype(np.float32)
            nxt = ay[:, y + 1] if y + 1 < r1 else np.zeros(O, np.float32)          # the last row closes the sum
            day = (ay[:, y] - nxt).astype(np.float32)
            for o in np.nonzero(day)[0]:
                cols = np.nonzero(ax[o])[0]
                if len(cols) == 0:
                    continue
                visits += 1
                dot = (C[:, cols] * ax[o, cols]).astype(np.float32).sum(axis=1, dtype=np.float32)
                part[o] = (part[o] + day[o] * dot).astype(np.float32)
        dv = (dv + part).astype(np.float32)
    return dv, visits


@pytest.mark.parametrize("rows_per_range", [16, 32, 64])
@pytest.mark.parametrize("same_sign", [False, True])
def test_summation_by_parts_matches_the_oracle_gradient(rows_per_range, same_sign):
    H = W = 64
    D = 6
    vocab = synth.Vocab(0)
    g = synth.make_graph(4242, 12, 24, vocab, include_dummies=False, mask_size=4)
    boxes = g.boxes.astype(np.float32)
    O = len(boxes)
    S, ay, ax = _separable_factors(boxes, H, W)
    assert np.abs(ay[:, :, None] * ax[:, None, :] - S).max() <= 2e-6              # the weights are separable
    dout = synth.det_tensor((D, H, W), 17, 1.0).astype(np.float32)
    if same_sign:
        dout = (1.0 + 0.25 * np.abs(dout)).astype(np.float32)
    vecs = torch.zeros((O, D), requires_grad=True)
    canvas = olayout.boxes_to_layout(vecs, torch.from_numpy(boxes), H, W)
    canvas.backward(torch.from_numpy(dout)[None])
    ref = vecs.grad.numpy()
    got, visits = _colsum_dvecs(dout, ay, ax, rows_per_range)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-5 * scale, (np.abs(got - ref).max() / scale)
    # the point of the exercise: an object is visited on the rows where its factor changes, not on every row it covers
    covered = int(sum(np.count_nonzero(ay[o]) for o in range(O) if np.count_nonzero(ax[o])))
    assert visits < 0.75 * covered, (visits, covered)
