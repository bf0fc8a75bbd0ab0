"""CUDA layout compositor (csrc/layout.cu through the C ABI) vs the reference's golden outputs and the
oracle.  Tolerance: 1e-5 relative to the tensor scale in fp32 (north_star)."""
import numpy as np
import pytest
import torch

from canonicalsg2im_b200 import synth
from tests import golden_inputs as gi
from tests.util import t, assert_close

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def L():
    from canonicalsg2im_b200 import layout
    return layout


def test_golden_demo(golden, L):
    g = golden("layout")
    v, b, m = t(g["demo_vecs"]), t(g["demo_boxes"]), t(g["demo_masks"])
    # D=3 is padded to 4 channels for the kernel (vector width); the extra channel is dropped
    v4 = torch.cat([v, torch.zeros(6, 1, device="cuda")], 1)
    assert_close(L.boxes_to_layout(v4, b, 64)[:, :3], g["demo_boxes64_out"], TOL, "boxes64")
    assert_close(L.masks_to_layout(v4, b, m, 64)[:, :3], g["demo_masks64_out"], TOL, "masks64")
    assert_close(L.boxes_to_layout(v4, b, 64, align_corners=True)[:, :3], g["demo_boxes64_legacy_out"], TOL, "legacy")
    assert_close(L.masks_to_layout(v4, b, m, 64, align_corners=True)[:, :3], g["demo_masks64_legacy_out"], TOL, "legacy m")


def test_golden_random_fwd_bwd(golden, L):
    g = golden("layout")
    for tag, masks in [("rnd_boxes", None), ("rnd_masks", g["rnd_masks"])]:
        v = t(g["rnd_vecs"]).requires_grad_(True)
        b = t(g["rnd_boxes"])
        if masks is None:
            y = L.boxes_to_layout(v, b, 32, 48)
        else:
            y = L.masks_to_layout(v, b, t(masks), 32, 48)
        assert y.shape == (1, 16, 32, 48)
        assert_close(y, g[tag + "_out"], TOL, tag)
        (y * t(gi.layout_out_grad(y.shape))).sum().backward()
        assert_close(v.grad, g[tag + "_dvecs"], TOL, tag + " dvecs")
    y = L.boxes_to_layout(t(g["rnd_vecs"]), t(g["rnd_boxes"]), 32, 48, pooling="avg")
    assert_close(y, g["rnd_boxes_avg_out"], TOL, "avg")
    # float masks
    v = t(g["rnd_vecs"]).requires_grad_(True)
    y = L.masks_to_layout(v, t(g["rnd_boxes"]), t(g["rnd_masks_f_in"]), 40, 40)
    assert_close(y, g["rnd_masks_f_out"], TOL, "float masks")
    (y * t(gi.layout_out_grad(y.shape))).sum().backward()
    assert_close(v.grad, g["rnd_masks_f_dvecs"], TOL, "float masks dvecs")


def _rand_objs(seed, n_img, n_min, n_max, D, M):
    vocab = synth.Vocab(0)
    vecs, boxes, masks, off = [], [], [], [0]
    for i in range(n_img):
        g = synth.make_graph(seed * 100 + i, n_min, n_max, vocab, include_dummies=False, mask_size=M)
        n = len(g.boxes)
        vecs.append(synth.det_tensor((n, D), seed * 1000 + i, 1.0))
        boxes.append(g.boxes)
        masks.append(g.masks)
        off.append(off[-1] + n)
    return np.concatenate(vecs), np.concatenate(boxes), np.concatenate(masks), np.array(off, np.int32)


@pytest.mark.parametrize("H,W,D,M,use_masks", [(64, 64, 128, 16, False), (64, 64, 128, 16, True),
                                               (40, 52, 8, 5, True), (17, 30, 4, 3, False),
                                               (128, 256, 32, 16, True)])
def test_batched_equals_per_image_oracle(L, H, W, D, M, use_masks):
    """Ragged launch == torch.cat of per-image reference calls (generator.py:81-96), fwd + dvecs."""
    from oracle import layout as olayout
    vecs, boxes, masks, off = _rand_objs(3, 5, 1, 9, D, M)
    v = t(vecs).requires_grad_(True)
    y = L.layout_batched(v, t(boxes), t(off), H, W, masks=t(masks) if use_masks else None, max_objs_per_image=9)
    gy = synth.det_tensor(tuple(y.shape), 91, 1.0)
    (y * t(gy)).sum().backward()
    vc = torch.from_numpy(vecs).requires_grad_(True)
    outs = olayout.batched_layout([vc[off[i]:off[i + 1]] for i in range(5)],
                                  [torch.from_numpy(boxes[off[i]:off[i + 1]]) for i in range(5)],
                                  [torch.from_numpy(masks[off[i]:off[i + 1]]) for i in range(5)] if use_masks else None,
                                  H, W)
    (outs * torch.from_numpy(gy)).sum().backward()
    assert_close(y, outs, TOL, "batched fwd")
    assert_close(v.grad, vc.grad, TOL, "batched dvecs")


@pytest.mark.parametrize("H,W,D,n_img,n_min,n_max,splits", [
    (128, 128, 32, 3, 1, 9, 0),      # 16-column strips
    (64, 256, 32, 2, 1, 9, 0),       # 32-column strips, one row per band
    (256, 256, 32, 2, 3, 20, 0),     # row ranges of a single row
    (64, 64, 64, 3, 40, 70, 0),      # more than 32 objects per image: several chunks
    (36, 64, 32, 2, 1, 9, 0),        # odd band count: one row range per image
    (64, 64, 32, 3, 3, 20, 1),       # the longest running sums: 64 rows per CTA (large batches select this)
    (64, 64, 32, 3, 3, 20, 2)])      # 32 rows per CTA (the cfg2 launch)
def test_column_sum_backward_shapes(L, monkeypatch, H, W, D, n_img, n_min, n_max, splits):
    """d/dvecs of boxes_to_layout on the shapes that select layout_bwd_colsum_kernel<8 / 16 / 32> (running column
    sums + summation by parts along y), against the per-image oracle: every strip width, several object chunks
    per image, an image without objects, row ranges from one row to the whole image (CSG_LAYOUT_SPLITS is read
    on every call)."""
    from oracle import layout as olayout
    if splits:
        monkeypatch.setenv("CSG_LAYOUT_SPLITS", str(splits))
    vecs, boxes, _, off = _rand_objs(11, n_img, n_min, n_max, D, 4)
    off = np.insert(off, 1, off[1]).astype(np.int32)                    # an empty image after the first one
    N = len(off) - 1
    v = t(vecs).requires_grad_(True)
    y = L.layout_batched(v, t(boxes), t(off), H, W, max_objs_per_image=int(np.diff(off).max()))
    gy = synth.det_tensor(tuple(y.shape), 92, 1.0)
    (y * t(gy)).sum().backward()
    vc = torch.from_numpy(vecs).requires_grad_(True)
    outs = olayout.batched_layout([vc[off[i]:off[i + 1]] for i in range(N)],
                                  [torch.from_numpy(boxes[off[i]:off[i + 1]]) for i in range(N)], None, H, W)
    (outs * torch.from_numpy(gy)).sum().backward()
    assert_close(y, outs, TOL, "fwd")
    assert_close(v.grad, vc.grad, TOL, "dvecs (column sums)")


def test_many_objects_chunking(L):
    """More objects per tile than the shared-memory list holds -> multi-pass accumulation."""
    from oracle import layout as olayout
    n = 70
    boxes = np.stack([synth.det_uniform(n, 1) * 0.3, synth.det_uniform(n, 2) * 0.3,
                      0.5 + 0.2 * synth.det_uniform(n, 3), 0.5 + 0.2 * synth.det_uniform(n, 4)], 1).astype(np.float32)
    vecs = synth.det_tensor((n, 8), 5, 1.0)
    v = t(vecs).requires_grad_(True)
    y = L.boxes_to_layout(v, t(boxes), 32, 64)
    vc = torch.from_numpy(vecs).requires_grad_(True)
    ref = olayout.boxes_to_layout(vc, torch.from_numpy(boxes), 32, 64)
    assert_close(y, ref, TOL, "70 objects fwd")
    gy = synth.det_tensor(tuple(y.shape), 6, 1.0)
    (y * t(gy)).sum().backward()
    (ref * torch.from_numpy(gy)).sum().backward()
    assert_close(v.grad, vc.grad, TOL, "70 objects dvecs")


def test_edge_cases(L):
    # zero objects: the reference raises (layout.py:128); defined here as an all-zero canvas
    y = L.layout_batched(torch.zeros(0, 8, device="cuda"), torch.zeros(0, 4, device="cuda"),
                         torch.tensor([0, 0, 0], dtype=torch.int32, device="cuda"), 16, 16)
    assert y.shape == (2, 8, 16, 16) and (y == 0).all()
    # degenerate box poisons the whole canvas with NaN, like grid_sample does
    v = torch.ones(2, 4, device="cuda")
    b = torch.tensor([[0.1, 0.1, 0.5, 0.5], [0.2, 0.2, 0.0, 0.3]], device="cuda")
    y = L.boxes_to_layout(v, b, 16)
    assert torch.isnan(y).all()
    # padded (-1) boxes and boxes outside the canvas contribute exact zeros
    b = torch.tensor([[-1., -1., -1., -1.], [1.5, 1.5, 0.2, 0.2]], device="cuda")
    assert (L.boxes_to_layout(v, b, 16) == 0).all()
    with pytest.raises(ValueError):
        L.boxes_to_layout(v, b, 16, pooling="max")


def test_full_size_properties(L):
    """cfg3 size (B=16, 256x256, D=128, M=16): linearity in vecs and additivity over objects — the
    size-independent properties of the compositor; no oracle needed at 512 MiB."""
    vecs, boxes, masks, off = _rand_objs(11, 16, 3, 8, 128, 16)
    v, b, m, o = t(vecs), t(boxes), t(masks), t(off)
    y1 = L.layout_batched(v, b, o, 256, 256, masks=m, max_objs_per_image=8)
    assert y1.shape == (16, 128, 256, 256)
    y2 = L.layout_batched(2.5 * v, b, o, 256, 256, masks=m, max_objs_per_image=8)
    assert_close(y2, 2.5 * y1, 1e-6, "linearity")
    del y2
    # additivity: splitting every image's objects into two calls and summing gives the same canvas
    keep = torch.arange(v.shape[0], device="cuda") % 2 == 0
    ya = L.layout_batched(torch.where(keep[:, None], v, torch.zeros_like(v)), b, o, 256, 256, masks=m)
    ya += L.layout_batched(torch.where(keep[:, None], torch.zeros_like(v), v), b, o, 256, 256, masks=m)
    assert_close(ya, y1, 1e-6, "additivity")
    # adjoint identity <y, G> == <v, dvecs(G)> ties the backward kernel to the forward one
    vg = v.clone().requires_grad_(True)
    y = L.layout_batched(vg, b, o, 256, 256, masks=m, max_objs_per_image=8)
    G = torch.randn(y.shape, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    lhs = (y.double() * G.double()).sum()
    y.backward(G)
    rhs = (vg.detach().double() * vg.grad.double()).sum()
    assert abs(lhs.item() - rhs.item()) <= 1e-5 * abs(lhs.item())


# ------------------------------------------------------------------------------------------------------
# test-mode occlusion compositor (layout.py:72-76, :135-147)
# ------------------------------------------------------------------------------------------------------
def test_golden_occlusion(golden, L):
    g = golden("layout")
    v, b, m = t(g["demo_vecs"]), t(g["demo_boxes"]), t(g["demo_masks"])
    v4 = torch.cat([v, torch.zeros(6, 1, device="cuda")], 1)
    y = L.masks_to_layout(v4, b, m, 64, test_mode=True)
    assert_close(y[:, :3], g["demo_masks64_test_out"], TOL, "demo occlusion")
    assert (y[:, 3] == 0).all()
    y = L.masks_to_layout(t(g["rnd_vecs"]), t(g["rnd_boxes"]), t(g["rnd_masks"]), 32, 48, test_mode=True)
    assert_close(y, g["rnd_masks_test_out"], TOL, "rnd occlusion")


@pytest.mark.parametrize("H,W,D,M,nmax", [(64, 64, 32, 16, 9), (40, 52, 8, 5, 6), (128, 128, 16, 16, 40)])
def test_occlusion_batched_vs_oracle(L, H, W, D, M, nmax):
    """Ragged occlusion launch == per-image oracle calls; nmax=40 also exercises more objects than one list chunk
    when they all overlap the tile."""
    from oracle import layout as olayout
    vecs, boxes, masks, off = _rand_objs(5, 4, 2, nmax, D, M)
    y = L.layout_batched(t(vecs), t(boxes), t(off), H, W, masks=t(masks), test_mode=True)
    ref = olayout.batched_layout([torch.from_numpy(vecs[off[i]:off[i + 1]]) for i in range(4)],
                                 [torch.from_numpy(boxes[off[i]:off[i + 1]]) for i in range(4)],
                                 [torch.from_numpy(masks[off[i]:off[i + 1]]) for i in range(4)], H, W, test_mode=True)
    # a pixel whose clean sample is within rounding of the 0.5 threshold may legitimately flip owner: allow a
    # handful of such pixels, everything else must agree to 1e-5
    diff = (y.cpu() - ref).abs().amax(dim=1)
    scale = ref.abs().max().item()
    bad = (diff > TOL * scale).sum().item()
    assert bad <= 2e-4 * diff.numel(), "occlusion: %d / %d pixels differ" % (bad, diff.numel())


def test_occlusion_properties_full_size(L):
    """cfg4-like size (10 images, 32-64 objects, 256x256): every pixel is either 0 or exactly one object's
    vec * weight with weight > 0.5, and the canvas is invariant to the storage order of the objects."""
    vecs, boxes, masks, off = _rand_objs(21, 10, 32, 64, 32, 16)
    v, b, m, o = t(vecs), t(boxes), t(masks).float(), t(off)
    y = L.layout_batched(v, b, o, 256, 256, masks=m, test_mode=True)
    assert y.shape == (10, 32, 256, 256) and torch.isfinite(y).all()
    # permute objects inside every image: masses are distinct, so the owner of each pixel is unchanged
    perm = np.concatenate([off[i] + np.random.RandomState(i).permutation(off[i + 1] - off[i]) for i in range(10)])
    pt = torch.from_numpy(perm).cuda()
    y2 = L.layout_batched(v[pt], b[pt], o, 256, 256, masks=m[pt], test_mode=True)
    assert torch.equal(y, y2)
    # the summed compositor upper-bounds coverage: where the plain sum has no mass at all, occlusion is 0
    ys = L.layout_batched(v.abs(), b, o, 256, 256, masks=m)
    assert (y[(ys == 0)] == 0).all()


# ------------------------------------------------------------------------------------------------------
# gradients wrt boxes and float masks (autograd of grid_sample + _boxes_to_grid)
# ------------------------------------------------------------------------------------------------------
def _geom_case(L, g, tag, H, W, masks=None, float_masks=False, legacy=False, pad=0):
    v = t(g["demo_vecs"] if tag.startswith("demo") else g["rnd_vecs"])
    b = t(g["demo_boxes"] if tag.startswith("demo") else g["rnd_boxes"]).requires_grad_(True)
    if pad:
        v = torch.cat([v, torch.zeros(v.shape[0], pad, device="cuda")], 1)
    m = None
    if masks is not None:
        m = t(masks)
        if float_masks:
            m = m.float().requires_grad_(True)
    y = L.layout_batched(v, b, torch.tensor([0, v.shape[0]], dtype=torch.int32, device="cuda"), H, W, masks=m,
                         align_corners=legacy)
    D = y.shape[1] - pad
    gy = torch.zeros_like(y)
    gy[:, :D] = t(gi.layout_out_grad((1, D, H, W)))
    (y * gy).sum().backward()
    assert_close(b.grad, g[tag + "_dboxes"], TOL, tag + " dboxes")
    if float_masks:
        assert_close(m.grad, g[tag + "_dmasks"], TOL, tag + " dmasks")


def test_golden_box_and_mask_grads(golden, L):
    g = golden("layout")
    _geom_case(L, g, "demo_boxes64", 64, 64, pad=1)
    _geom_case(L, g, "demo_boxes64_legacy", 64, 64, legacy=True, pad=1)
    _geom_case(L, g, "demo_masks64", 64, 64, masks=g["demo_masks"], float_masks=True, pad=1)
    _geom_case(L, g, "demo_masks64_legacy", 64, 64, masks=g["demo_masks"], float_masks=True, legacy=True, pad=1)
    _geom_case(L, g, "rnd_boxes", 32, 48)
    _geom_case(L, g, "rnd_masks", 32, 48, masks=g["rnd_masks"])
    _geom_case(L, g, "rnd_masks_f", 40, 40, masks=g["rnd_masks_f_in"], float_masks=True)


def test_geom_grads_batched_vs_oracle(L):
    from oracle import layout as olayout
    H, W, D, M = 48, 64, 16, 8
    vecs, boxes, masks, off = _rand_objs(9, 3, 2, 7, D, M)
    fm = synth.det_uniform(masks.size, 3).reshape(masks.shape).astype(np.float32)
    b = t(boxes).requires_grad_(True)
    m = t(fm).requires_grad_(True)
    y = L.layout_batched(t(vecs), b, t(off), H, W, masks=m)
    gy = synth.det_tensor(tuple(y.shape), 17, 1.0)
    (y * t(gy)).sum().backward()
    bc = torch.from_numpy(boxes).requires_grad_(True)
    mc = torch.from_numpy(fm).requires_grad_(True)
    ref = olayout.batched_layout([torch.from_numpy(vecs[off[i]:off[i + 1]]) for i in range(3)],
                                 [bc[off[i]:off[i + 1]] for i in range(3)],
                                 [mc[off[i]:off[i + 1]] for i in range(3)], H, W)
    (ref * torch.from_numpy(gy)).sum().backward()
    assert_close(b.grad, bc.grad, TOL, "batched dboxes")
    assert_close(m.grad, mc.grad, TOL, "batched dmasks")


def test_tensor_core_backward_path_matches_ring_path():
    """layout_bwd_tc.cu (tcgen05 kind::tf32, 3-term split) is selected with CSG_LAYOUT_BWD=tc (read once per process,
    hence the subprocess): its d/dvecs must agree with the default ring kernel to the fp32 contract (1e-5) for boxes
    and masks, including an image without objects."""
    import os, subprocess, sys, textwrap
    code = textwrap.dedent("""
        import sys, numpy as np, torch
        sys.path.insert(0, %r)
        from canonicalsg2im_b200 import synth
        from canonicalsg2im_b200.layout import layout_batched
        vocab = synth.Vocab(0)
        gs = [synth.make_graph(900 + i, 3, 20, vocab, include_dummies=False, mask_size=16) for i in range(9)]
        boxes = np.concatenate([g.boxes for g in gs]); masks = np.concatenate([g.masks for g in gs]).astype(np.float32)
        off = np.concatenate([[0], np.cumsum([len(g.boxes) for g in gs])]).astype(np.int32)
        off = np.insert(off, 4, off[4])                      # an empty image in the middle
        N, mo = len(off) - 1, int(np.diff(off).max())
        out = {}
        for use_masks in (False, True):
            v = torch.from_numpy(synth.det_tensor((len(boxes), 128), 5, 1.0)).cuda().requires_grad_(True)
            y = layout_batched(v, torch.from_numpy(boxes).cuda(), torch.from_numpy(off).cuda(), 64, 64,
                               masks=torch.from_numpy(masks).cuda() if use_masks else None, max_objs_per_image=mo)
            g = torch.from_numpy(synth.det_tensor(tuple(y.shape), 6, 1.0)).cuda()
            y.backward(g)
            out[use_masks] = v.grad.cpu().numpy()
        np.savez(sys.argv[1], boxes=out[False], masks=out[True])
    """ % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import tempfile
    res = {}
    with tempfile.TemporaryDirectory() as td:
        for mode in ("ring", "tc"):
            path = os.path.join(td, mode + ".npz")
            env = dict(os.environ, CSG_LAYOUT_BWD=mode)
            subprocess.run([sys.executable, "-c", code, path], check=True, env=env, timeout=300)
            res[mode] = dict(np.load(path))
    for k in ("boxes", "masks"):
        a, b = res["tc"][k], res["ring"][k]
        assert np.isfinite(a).all()
        assert np.abs(a - b).max() <= 1e-5 * np.abs(b).max(), k


# ------------------------------------------------------------------------------------------------------
# SPADE-side pyramid: generator.py:99 + normalization.py:102
# ------------------------------------------------------------------------------------------------------
def test_layout_pyramid_vs_reference_interpolate(golden, L):
    """layout_pyramid's levels against the reference's own F.interpolate(mode='nearest') of the reference canvas
    (tests/golden/pyramid.npz, unmodified reference), 1e-5; and bit-identical to slicing our own canvas."""
    g = golden("pyramid")
    vocab = synth.Vocab(0)
    graphs = synth.make_graphs(3, 31, 2, 6, vocab, include_dummies=False)
    vecs = np.concatenate([synth.det_tensor((len(gr.boxes), 16), 50 + i, 1.0) for i, gr in enumerate(graphs)])
    boxes = np.concatenate([gr.boxes for gr in graphs])
    off = np.concatenate([[0], np.cumsum([len(gr.boxes) for gr in graphs])]).astype(np.int32)
    canvas, levels = L.layout_pyramid(t(vecs), t(boxes), t(off), 64, 64, sizes=[(2, 2), 4, 8, 16, 32])
    assert_close(canvas, g["seg"], TOL, "canvas")
    for lev, name, r in zip(levels, ["head", "l4", "l8", "l16", "l32"], [32, 16, 8, 4, 2]):
        assert_close(lev, g[name], TOL, "pyramid " + name)
        assert torch.equal(lev, canvas[:, :, ::r, ::r])
    # a non-integer ratio follows ATen's source-index rule as well
    _, (l24,) = L.layout_pyramid(t(vecs), t(boxes), t(off), 64, 64, sizes=[24])
    ref24 = torch.nn.functional.interpolate(canvas, size=(24, 24), mode="nearest")
    assert torch.equal(l24, ref24)


def test_layout_pyramid_gradients_vs_oracle(L):
    """d vecs / d boxes through canvas + levels == autograd of the oracle canvas followed by F.interpolate."""
    from oracle import layout as olayout
    import torch.nn.functional as F
    vecs, boxes, masks, off = _rand_objs(17, 3, 2, 7, 16, 8)
    sizes = [4, 16, 32]
    v, b = t(vecs).requires_grad_(True), t(boxes).requires_grad_(True)
    canvas, levels = L.layout_pyramid(v, b, t(off), 64, 64, sizes=sizes, masks=t(masks), max_objs_per_image=7)
    gs = [synth.det_tensor(tuple(x.shape), 200 + i, 1.0) for i, x in enumerate([canvas] + levels)]
    sum((x * t(gg)).sum() for x, gg in zip([canvas] + levels, gs)).backward()
    vc, bcx = torch.from_numpy(vecs).requires_grad_(True), torch.from_numpy(boxes).requires_grad_(True)
    ref = olayout.batched_layout([vc[off[i]:off[i + 1]] for i in range(3)], [bcx[off[i]:off[i + 1]] for i in range(3)],
                                 [torch.from_numpy(masks[off[i]:off[i + 1]]) for i in range(3)], 64, 64)
    rl = [F.interpolate(ref, size=(s_, s_), mode="nearest") for s_ in sizes]
    sum((x * torch.from_numpy(gg)).sum() for x, gg in zip([ref] + rl, gs)).backward()
    assert_close(v.grad, vc.grad, TOL, "pyramid dvecs")
    assert_close(b.grad, bcx.grad, TOL, "pyramid dboxes")
