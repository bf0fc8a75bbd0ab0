"""Oracle (torch CPU) for the layout compositor and the bilinear crop.
TEST INFRASTRUCTURE ONLY.

Restates ``sg2im/layout.py:12-188`` and ``sg2im/bilinear.py:44-94,155-184``
of the reference.  The arithmetic lives in ``torch.linspace`` and
``F.grid_sample`` (bilinear, zero padding); the reference calls grid_sample
without ``align_corners`` which under torch 2.11 means ``False`` (the value the
golden vectors were generated with).  ``align_corners`` is exposed so that the
pre-1.3 behaviour the checkpoints were trained with can be checked as well.

Boxes are [x0, y0, w, h] (layout.py:95-96), despite the reference docstrings.
"""
import numpy as np
import torch
import torch.nn.functional as F


def boxes_to_grid(boxes, H, W):
    """sg2im/layout.py:80-112 -> [O, H, W, 2] sampling grid in [-1, 1]."""
    O = boxes.shape[0]
    b = boxes.view(O, 4, 1, 1)
    x0, y0, ww, hh = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    X = torch.linspace(0, 1, steps=W).view(1, 1, W).to(boxes)                # fp32 linspace, :98
    Y = torch.linspace(0, 1, steps=H).view(1, H, 1).to(boxes)                # :99
    X = ((X - x0) / ww).expand(O, H, W)                                      # :101,105
    Y = ((Y - y0) / hh).expand(O, H, W)                                      # :102,106
    return torch.stack([X, Y], dim=3).mul(2).sub(1)                          # :107-110


def pool_samples(samples, pooling="sum"):
    """sg2im/layout.py:156-188 — all objects belong to image 0; 'avg' divides
    by the object count (the reference also prints the counts)."""
    O = samples.shape[0]
    out = torch.zeros((1,) + tuple(samples.shape[1:]), dtype=samples.dtype)
    idx = torch.zeros_like(samples, dtype=torch.long)
    out = out.scatter_add(0, idx, samples)                                   # :172-174
    if pooling == "avg":
        out = out / max(O, 1)                                                # :176-184
    elif pooling != "sum":
        raise ValueError('Invalid pooling "%s"' % pooling)
    return out


def boxes_to_layout(vecs, boxes, H, W=None, pooling="sum", align_corners=False):
    """sg2im/layout.py:12-45."""
    O, D = vecs.shape
    W = H if W is None else W
    grid = boxes_to_grid(boxes, H, W)
    img_in = vecs.view(O, D, 1, 1).expand(O, D, 8, 8)                        # :34
    sampled = F.grid_sample(img_in, grid, mode="bilinear", padding_mode="zeros",
                            align_corners=align_corners)                     # :35
    return pool_samples(sampled, pooling)


def pool_mask_samples(samples, clean, pooling="sum"):
    """sg2im/layout.py:115-153.  ``clean is None`` -> plain sum over objects;
    otherwise the test-mode occlusion compositor: objects in ascending order of
    total sampled mass, first writer wins where its clean mask > 0.5."""
    O = samples.shape[0]
    if clean is None:
        out = torch.zeros((1,) + tuple(samples.shape[1:]), dtype=samples.dtype)
        out = out.scatter_add(0, torch.zeros_like(samples, dtype=torch.long), samples)
    else:
        mass = [torch.sum(samples[j]).item() for j in range(O)]              # :139
        order = np.argsort(mass)                                             # :140
        result = torch.zeros(samples.shape[1:], dtype=samples.dtype)
        taken = torch.zeros(samples.shape[2:], dtype=samples.dtype)
        for j in order:                                                      # :143-146
            m = (taken == 0).float() * (clean[j, 0] > 0.5).float()
            taken += m
            result += samples[j] * m
        out = result[None]
    if pooling != "sum":
        raise ValueError('Invalid pooling "%s"' % pooling)
    return out


def masks_to_layout(vecs, boxes, masks, H, W=None, pooling="sum", test_mode=False,
                    align_corners=False):
    """sg2im/layout.py:48-77."""
    O, D = vecs.shape
    M = masks.shape[1]
    assert tuple(masks.shape) == (O, M, M)
    W = H if W is None else W
    grid = boxes_to_grid(boxes, H, W)
    img_in = vecs.view(O, D, 1, 1) * masks.float().view(O, 1, M, M)          # :69
    sampled = F.grid_sample(img_in, grid, mode="bilinear", padding_mode="zeros",
                            align_corners=align_corners)                     # :70
    clean = None
    if test_mode:
        clean = F.grid_sample(masks.float().view(O, 1, M, M), grid, mode="bilinear",
                              padding_mode="zeros", align_corners=align_corners)  # :73
    return pool_mask_samples(sampled, clean, pooling)


def batched_layout(vecs_list, boxes_list, masks_list, H, W=None, test_mode=False,
                   align_corners=False):
    """spade/models/networks/generator.py:81-96 — per-image loop + cat."""
    outs = []
    for i, (v, b) in enumerate(zip(vecs_list, boxes_list)):
        if masks_list is not None:
            outs.append(masks_to_layout(v, b, masks_list[i], H, W, test_mode=test_mode,
                                        align_corners=align_corners))
        else:
            outs.append(boxes_to_layout(v, b, H, W, align_corners=align_corners))
    return torch.cat(outs, dim=0)


def tensor_linspace(start, end, steps):
    """sg2im/bilinear.py:155-184 — start*linspace(1,0) + end*linspace(0,1)."""
    shape = tuple(start.shape) + (steps,)
    w0 = torch.linspace(1, 0, steps=steps).to(start).view((1,) * start.dim() + (steps,)).expand(shape)
    w1 = torch.linspace(0, 1, steps=steps).to(start).view((1,) * start.dim() + (steps,)).expand(shape)
    return w0 * start.contiguous().view(tuple(start.shape) + (1,)).expand(shape) + \
        w1 * end.contiguous().view(tuple(end.shape) + (1,)).expand(shape)


def crop_bbox(feats, bbox, HH, WW=None, align_corners=False):
    """sg2im/bilinear.py:65-94 (backend='cudnn'): feats [N,C,H,W], bbox [N,4] xywh
    -> crops [N,C,HH,WW]."""
    WW = HH if WW is None else WW
    N = feats.shape[0]
    pts = bbox.clone()                                                       # sg2im/metrics.py:4-8
    pts[:, 2] = bbox[:, 0] + bbox[:, 2]
    pts[:, 3] = bbox[:, 1] + bbox[:, 3]
    pts = 2 * pts - 1                                                        # :84
    X = tensor_linspace(pts[:, 0], pts[:, 2], WW).view(N, 1, WW).expand(N, HH, WW)
    Y = tensor_linspace(pts[:, 1], pts[:, 3], HH).view(N, HH, 1).expand(N, HH, WW)
    return F.grid_sample(feats, torch.stack([X, Y], dim=3), mode="bilinear",
                         padding_mode="zeros", align_corners=align_corners)  # :92-94


def crop_bbox_batch(imgs, obj_keep, bbox, HH, WW=None, align_corners=False):
    """sg2im/bilinear.py:44-62: imgs [N,C,H,W], obj_keep [N,O] bool (the
    ``remove_dummy_objects`` mask, sg2im/utils.py:56-63), bbox [N,O,4]."""
    N, C, H, W = imgs.shape
    feats, boxes = [], []
    for i in range(N):
        cur = bbox[i][obj_keep[i]]
        feats.append(imgs[i].view(1, C, H, W).expand(cur.shape[0], C, H, W).contiguous())
        boxes.append(cur)
    return crop_bbox(torch.cat(feats, 0), torch.cat(boxes, 0), HH, WW, align_corners)


def bilinear_sample(feats, X, Y):
    """sg2im/bilinear.py:97-152: four clamped taps (floor, floor + 1) with weights measured against the clamped taps.
    feats [N,C,H,W]; X, Y [N,HH,WW] in [0, 1]."""
    N, C, H, W = feats.shape
    _, HH, WW = X.shape
    X = X * W                                                                # :114-115
    Y = Y * H
    x0 = X.floor().clamp(min=0, max=W - 1)                                   # :118-121
    x1 = (x0 + 1).clamp(min=0, max=W - 1)
    y0 = Y.floor().clamp(min=0, max=H - 1)
    y1 = (y0 + 1).clamp(min=0, max=H - 1)
    flat = feats.reshape(N, C, H * W)

    def take(yy, xx):                                                        # :131-143
        idx = (W * yy + xx).view(N, 1, HH * WW).expand(N, C, HH * WW).long()
        return flat.gather(2, idx).view(N, C, HH, WW)
    v1, v2, v3, v4 = take(y0, x0), take(y1, x0), take(y0, x1), take(y1, x1)
    e = lambda w: w.view(N, 1, HH, WW).expand(N, C, HH, WW)                  # :146-149
    w1, w2 = e((x1 - X) * (y1 - Y)), e((x1 - X) * (Y - y0))
    w3, w4 = e((X - x0) * (y1 - Y)), e((X - x0) * (Y - y0))
    return w1 * v1 + w2 * v2 + w3 * v3 + w4 * v4                             # :152


def crop_bbox_jj(feats, bbox, HH, WW=None):
    """sg2im/bilinear.py:65-94 with backend='jj': the box stays in [0, 1] coordinates and is sampled by
    ``bilinear_sample``."""
    WW = HH if WW is None else WW
    N = feats.shape[0]
    pts = bbox.clone()
    pts[:, 2] = bbox[:, 0] + bbox[:, 2]
    pts[:, 3] = bbox[:, 1] + bbox[:, 3]
    X = tensor_linspace(pts[:, 0], pts[:, 2], WW).view(N, 1, WW).expand(N, HH, WW)
    Y = tensor_linspace(pts[:, 1], pts[:, 3], HH).view(N, HH, 1).expand(N, HH, WW)
    return bilinear_sample(feats, X, Y)
