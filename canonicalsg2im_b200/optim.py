"""Adam for the scene-graph -> layout model on one multi-tensor kernel (``csrc/optim.cu``).

The reference trains with ``torch.optim.Adam`` (``scripts/train.py``); this class keeps its constructor arguments,
``step`` / ``zero_grad`` and per-parameter state (``step``, ``exp_avg``, ``exp_avg_sq``) and performs the same
arithmetic, but updates all parameters in ceil(#tensors / 48) launches with every element read and written once.
CUDA fp32 parameters only; there is no CPU path.
"""
import ctypes

import torch

from . import _lib
from .ops import lib, _stream


class FusedAdam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.CsgError("FusedAdam needs contiguous fp32 CUDA parameters (there is no CPU path)")
        self.lr, self.betas, self.eps, self.weight_decay = lr, tuple(betas), eps, weight_decay
        self.state = {}
        self.t = 0

    def _state(self, p):
        st = self.state.get(p)
        if st is None:
            st = self.state[p] = {"exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        return st

    @torch.no_grad()
    def step(self):
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        self.t += 1
        n = len(live)
        grads = []
        for p in live:
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            grads.append(g)
        VP, IA = ctypes.c_void_p * n, ctypes.c_int * n
        sts = [self._state(p) for p in live]
        rc = lib().csg_adam_multi(n, VP(*[p.data_ptr() for p in live]), VP(*[g.data_ptr() for g in grads]),
                                  VP(*[s["exp_avg"].data_ptr() for s in sts]),
                                  VP(*[s["exp_avg_sq"].data_ptr() for s in sts]), IA(*[p.numel() for p in live]),
                                  float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                  float(self.weight_decay), self.t, _stream())
        _lib.check(rc, "csg_adam_multi")

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def state_dict(self):
        return {"t": self.t, "lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                "state": [{k: v.clone() for k, v in self._state(p).items()} for p in self.params]}

    def load_state_dict(self, sd):
        self.t, self.lr, self.betas, self.eps = sd["t"], sd["lr"], tuple(sd["betas"]), sd["eps"]
        self.weight_decay = sd["weight_decay"]
        for p, st in zip(self.params, sd["state"]):
            mine = self._state(p)
            for k in mine:
                mine[k].copy_(st[k])
