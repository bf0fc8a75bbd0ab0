"""Differentiable box crops: drop-in for ``sg2im/bilinear.py`` of the reference.

``crop_bbox(feats, bbox, HH, WW=None, backend='cudnn')`` (``bilinear.py:65-94``) and
``crop_bbox_batch(imgs, objs, bbox, HH, WW=None, vocab=None, backend='cudnn')``
(``bilinear.py:13-62``) keep the reference's signatures; ``crop_bbox_ragged`` is the form the kernels
implement (crops ``crop_off[n]:crop_off[n+1]`` sample image ``n`` in place, no per-object image copies).
Gradients flow to ``feats`` (what the discriminator's generator loss needs, ``discriminator.py:249``); boxes
are ground truth on every caller of the reference and get no gradient.

All arithmetic runs in ``csg_crop_bbox_*`` (csrc/crop.cu); CPU tensors or a missing library raise.
"""
import torch

from . import _lib
from .ops import lib, ptr, need_cuda, f32c, workspace, _stream

_LIN = {}


def _lin_pair(steps, device):
    """``tensor_linspace`` weights (bilinear.py:173-176): linspace(1, 0) and linspace(0, 1) in fp32,
    built on the CPU and moved, like ``torch.linspace(...).to(start)``."""
    key = (int(steps), str(device))
    t = _LIN.get(key)
    if t is None:
        t = (torch.linspace(1, 0, steps=int(steps)).to(device), torch.linspace(0, 1, steps=int(steps)).to(device))
        _LIN[key] = t
    return t


class _CropFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, bbox, crop_off, HH, WW, align_corners):
        need_cuda(feats, bbox, crop_off)
        feats_c, bbox_c = f32c(feats), f32c(bbox)
        N, C, H, W = feats_c.shape
        NC = bbox_c.shape[0]
        swx, ewx = _lin_pair(WW, feats.device)
        swy, ewy = _lin_pair(HH, feats.device)
        out = torch.empty((NC, C, HH, WW), dtype=torch.float32, device=feats.device)
        L = lib()
        ws = workspace(L.csg_crop_bbox_workspace(NC), feats.device)
        rc = L.csg_crop_bbox_fwd(ptr(feats_c), ptr(bbox_c), ptr(crop_off), ptr(swx), ptr(ewx), ptr(swy), ptr(ewy),
                                 ptr(out), N, NC, C, H, W, HH, WW, int(align_corners), ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "csg_crop_bbox_fwd")
        ctx.save_for_backward(bbox_c, crop_off)
        ctx.dims = (N, NC, C, H, W, HH, WW, int(align_corners))
        return out

    @staticmethod
    def backward(ctx, dcrops):
        bbox, crop_off = ctx.saved_tensors
        N, NC, C, H, W, HH, WW, align = ctx.dims
        if ctx.needs_input_grad[1]:
            raise NotImplementedError("crop_bbox: gradients wrt bbox are not provided (every reference caller "
                                      "passes ground-truth boxes)")
        dfeats = None
        if ctx.needs_input_grad[0]:
            dcrops = f32c(dcrops)
            swx, ewx = _lin_pair(WW, dcrops.device)
            swy, ewy = _lin_pair(HH, dcrops.device)
            dfeats = torch.empty((N, C, H, W), dtype=torch.float32, device=dcrops.device)
            rc = lib().csg_crop_bbox_bwd(ptr(dcrops), ptr(bbox), ptr(crop_off), ptr(swx), ptr(ewx), ptr(swy), ptr(ewy),
                                         ptr(dfeats), N, NC, C, H, W, HH, WW, align, _stream())
            _lib.check(rc, "csg_crop_bbox_bwd")
        return dfeats, None, None, None, None, None


def crop_bbox_ragged(imgs, bbox, crop_off, HH, WW=None, align_corners=False):
    """imgs [N, C, H, W], bbox [NC, 4] xywh, crop_off [N+1] int32 -> crops [NC, C, HH, WW]."""
    WW = HH if WW is None else WW
    if crop_off.dtype != torch.int32:
        crop_off = crop_off.to(torch.int32)
    return _CropFn.apply(imgs, bbox, crop_off.to(imgs.device).contiguous(), int(HH), int(WW), bool(align_corners))


def crop_bbox(feats, bbox, HH, WW=None, backend="cudnn", align_corners=False):
    """``sg2im/bilinear.py:65-94``: feats [N, C, H, W], bbox [N, 4] xywh -> [N, C, HH, WW] (crop i of image i)."""
    N = feats.size(0)
    assert bbox.size(0) == N
    assert bbox.size(1) == 4
    if backend not in ("cudnn", "jj"):
        raise ValueError('Invalid backend "%s"' % backend)
    off = torch.arange(N + 1, dtype=torch.int32, device=feats.device)
    if backend == "jj":          # bilinear_sample (bilinear.py:97-152): the kernels' third coordinate mode
        return _CropFn.apply(feats, bbox, off, int(HH), int(HH if WW is None else WW), 2)
    return crop_bbox_ragged(feats, bbox, off, HH, WW, align_corners)


def crop_bbox_batch(imgs, objs, bbox, HH, WW=None, vocab=None, backend="cudnn", align_corners=False):
    """``sg2im/bilinear.py:13-62``: imgs [N, C, H, W], objs [N, O, A] int64, bbox [N, O, 4]; crops of the real
    objects (not padding, not ``__image__``: ``remove_dummy_objects``, sg2im/utils.py:56-63), image-major."""
    need_cuda(imgs, objs, bbox)
    N, O = objs.shape[0], objs.shape[1]
    first = objs.reshape(N, O, -1)[:, :, 0]
    keep = (first != 0) & (first != vocab["object_name_to_idx"]["__image__"])
    counts = keep.sum(dim=1)
    off = torch.zeros(N + 1, dtype=torch.int32, device=imgs.device)
    off[1:] = counts.cumsum(0)
    flat = bbox.reshape(N * O, 4)[keep.reshape(-1)]
    return crop_bbox_ragged(imgs, flat, off, HH, WW, align_corners)
