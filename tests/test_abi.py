"""Host-side checks that run without a GPU: the C-ABI library builds, loads and exports exactly what
include/csg2im.h declares; the product path refuses CPU tensors instead of falling back."""
import ctypes
import os
import subprocess

import pytest
import torch

from canonicalsg2im_b200 import _lib


@pytest.fixture(scope="module")
def built():
    _lib.build()
    return _lib.load()


def test_header_symbols_exported(built):
    protos = _lib.parse_header()
    assert len(protos) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in protos:
        assert hasattr(raw, name), "declared in include/csg2im.h but not exported: %s" % name


def test_no_undeclared_exports(built):
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("csg_")}
    assert exported == set(_lib.parse_header()), exported ^ set(_lib.parse_header())


def test_version_and_error_channel(built):
    assert built.csg_version() >= 100
    # a size error is reported through the return code + csg_last_error, without touching the GPU
    rc = built.csg_layout_fwd(0, 0, 0, 0, 0, 0, 0, 1, 6, 8, 8, 0, 0, 0, 0)     # D=6 is not a multiple of 4
    assert rc != 0 and b"multiple of 4" in built.csg_last_error()
    with pytest.raises(_lib.CsgError):
        _lib.check(rc, "csg_layout_fwd")


def test_sass_is_sm100(built):
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out, out


def test_cpu_tensors_are_refused(built):
    from canonicalsg2im_b200 import boxes_to_layout, GraphTripleConv
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        boxes_to_layout(torch.zeros(2, 8), torch.rand(2, 4), 16)
    layer = GraphTripleConv(8, 8, 8, 8, 16, 1, predicates_transitive_weights=torch.nn.Parameter(torch.zeros(4)))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        layer(torch.zeros(1, 3, 8), torch.zeros(1, 2, 8), torch.zeros(1, 2, 2, dtype=torch.long),
              torch.ones(1, 2, dtype=torch.bool), torch.zeros(1, 2, dtype=torch.long), torch.zeros(1, 2, dtype=torch.long))


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU/PyTorch fallback"):
        _lib.load()


def test_state_dict_keys_match_reference():
    """Drop-in contract: parameter names of the reference's modules."""
    from canonicalsg2im_b200.graph import GraphTripleConv
    w = torch.nn.Parameter(torch.zeros(8))
    layer = GraphTripleConv(128, 128, 128, 128, 512, 1, predicates_transitive_weights=w)
    keys = set(layer.state_dict().keys())
    assert keys == {"net1.0.weight", "net1.0.bias", "net1.2.weight", "net1.2.bias", "net2.0.weight", "net2.0.bias",
                    "net2.2.weight", "net2.2.bias", "predicates_transitive_weights"}
    assert layer.net1[0].weight.shape == (512, 384) and layer.net1[2].weight.shape == (1152, 512)
    assert layer.net2[0].weight.shape == (512, 512) and layer.net2[2].weight.shape == (128, 512)
    with pytest.raises(AssertionError):
        GraphTripleConv(8, 8, 8, 8, 8, 1, pooling="max")
