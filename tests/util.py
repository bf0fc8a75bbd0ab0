import numpy as np
import torch


def t(x, device="cuda"):
    return torch.from_numpy(np.ascontiguousarray(x)).to(device)


def rel_err(a, b):
    """max |a-b| relative to the scale of the reference tensor b (the tolerance form used throughout)."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu() if torch.is_tensor(b) else torch.from_numpy(np.asarray(b)).double()
    scale = max(b.abs().max().item(), 1e-30)
    return (a - b).abs().max().item() / scale


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, "%s: relative error %.3e > %.1e" % (what, e, tol)
