"""CUDA box crops (csrc/crop.cu through the C ABI) vs the reference's golden outputs and the oracle.
Tolerance 1e-5 relative to the tensor scale (fp32, north_star)."""
import numpy as np
import pytest
import torch

from canonicalsg2im_b200 import synth
from tests import golden_inputs as gi
from tests.util import t, assert_close

pytestmark = pytest.mark.gpu
TOL = 1e-5


def test_golden_crop_bbox(golden):
    from canonicalsg2im_b200.bilinear import crop_bbox
    g = golden("layout")
    im = t(g["crop_imgs"]).requires_grad_(True)
    crops = crop_bbox(im, t(g["crop_boxes"]), 8, 12)
    assert crops.shape == (3, 3, 8, 12)
    assert_close(crops, g["crop_out"], TOL, "crop fwd")
    (crops * t(gi.crop_out_grad(crops.shape))).sum().backward()
    assert_close(im.grad, g["crop_dimgs"], TOL, "crop dimgs")


@pytest.mark.parametrize("N,C,H,W,HH,WW", [(4, 3, 64, 64, 32, 32), (2, 5, 37, 50, 16, 24), (3, 3, 256, 256, 32, 32)])
def test_crop_bbox_batch_vs_oracle(N, C, H, W, HH, WW):
    """crop_bbox_batch (padding and __image__ rows dropped, images read in place) == the reference's
    expand-per-object + grid_sample (bilinear.py:44-62), forward and d/dimgs."""
    from canonicalsg2im_b200.bilinear import crop_bbox_batch
    from oracle import layout as olayout
    O = 6
    vocab = {"object_name_to_idx": {"__image__": 7}}
    imgs = synth.det_tensor((N, C, H, W), 51, 1.0)
    objs = synth.det_int(N * O, 52, 0, 9).reshape(N, O, 1)
    objs[0, 1, 0], objs[1, 0, 0], objs[N - 1, :, 0] = 7, 0, 0      # a dummy, a pad, an image with no real object
    objs[N - 1, 2, 0] = 3
    u = synth.det_uniform(N * O * 4, 53).reshape(N, O, 4).astype(np.float32)
    bbox = np.concatenate([u[..., :2] * 0.7 - 0.1, 0.05 + u[..., 2:] * 0.6], -1).astype(np.float32)
    im = t(imgs).requires_grad_(True)
    crops = crop_bbox_batch(im, t(objs), t(bbox), HH, WW, vocab=vocab)
    keep = (objs[:, :, 0] != 0) & (objs[:, :, 0] != 7)
    ic = torch.from_numpy(imgs).requires_grad_(True)
    ref = olayout.crop_bbox_batch(ic, torch.from_numpy(keep), torch.from_numpy(bbox), HH, WW)
    assert crops.shape == ref.shape
    assert_close(crops, ref, TOL, "crop batch fwd")
    gc = synth.det_tensor(tuple(ref.shape), 54, 1.0)
    (crops * t(gc)).sum().backward()
    (ref * torch.from_numpy(gc)).sum().backward()
    assert_close(im.grad, ic.grad, TOL, "crop batch dimgs")


def test_crop_identity_and_errors():
    from canonicalsg2im_b200.bilinear import crop_bbox
    # align_corners=True with the full-image box and HH=H is the identity resampling
    x = torch.randn(2, 3, 16, 20, device="cuda")
    full = torch.tensor([[0., 0., 1., 1.]] * 2, device="cuda")
    assert_close(crop_bbox(x, full, 16, 20, align_corners=True), x, 1e-5, "identity")
    with pytest.raises(RuntimeError):
        crop_bbox(x.cpu(), full.cpu(), 8)
    with pytest.raises(AssertionError):
        crop_bbox(x, full[:1], 8)


def test_crop_bbox_jj_backend_golden(golden):
    """backend='jj' (bilinear_sample, bilinear.py:97-152) against the unmodified reference: forward and d/dfeats."""
    from canonicalsg2im_b200.bilinear import crop_bbox
    g = golden("crop_jj")
    im = t(g["imgs"]).requires_grad_(True)
    crops = crop_bbox(im, t(g["boxes"]), 8, 12, backend="jj")
    assert_close(crops, g["out"], 1e-5, "crop jj")
    (crops * t(gi.crop_out_grad(crops.shape))).sum().backward()
    assert_close(im.grad, g["dimgs"], 1e-5, "crop jj d/dfeats")
