"""Adam for the scene-graph -> layout model on one multi-tensor kernel (``csrc/optim.cu``).

The reference trains with ``torch.optim.Adam`` (``scripts/train.py``); this class keeps its constructor arguments,
``step`` / ``zero_grad``, per-parameter state (``step``, ``exp_avg``, ``exp_avg_sq``) and ``state_dict`` layout and
performs the same arithmetic, but updates all parameters in ceil(#tensors / 48) launches with every element read and written once.
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
        self.post_step = []       # callables run after every step (e.g. Sg2LayoutModel.refresh_weight_copies)

    def _state(self, p):
        st = self.state.get(p)
        if st is None:
            # "step" counts the updates of THIS parameter (a parameter without .grad is skipped and its count does
            # not advance), exactly as torch.optim.Adam's per-parameter state
            st = self.state[p] = {"step": 0, "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
        return st

    @torch.no_grad()
    def step(self):
        live = [p for p in self.params if p.grad is not None]
        if not live:
            return
        n = len(live)
        grads = []
        for p in live:
            g = p.grad
            if g.dtype != torch.float32 or not g.is_contiguous():
                g = g.float().contiguous()
            grads.append(g)
        VP, IA = ctypes.c_void_p * n, ctypes.c_int * n
        sts = [self._state(p) for p in live]
        for s in sts:
            s["step"] += 1
        rc = lib().csg_adam_multi(n, VP(*[p.data_ptr() for p in live]), VP(*[g.data_ptr() for g in grads]),
                                  VP(*[s["exp_avg"].data_ptr() for s in sts]),
                                  VP(*[s["exp_avg_sq"].data_ptr() for s in sts]), IA(*[p.numel() for p in live]),
                                  float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps),
                                  float(self.weight_decay), IA(*[s["step"] for s in sts]), _stream())
        _lib.check(rc, "csg_adam_multi")
        # the kernel wrote the parameters through raw pointers: autograd's version counters did not move, so the cached
        # 16-bit weight copies of the tensor-core engine (graph_tc._WCOPIES) must be dropped here ...
        from . import graph_tc
        graph_tc.invalidate_weight_copies(live)
        for cb in self.post_step:       # ... and are rebuilt in one launch by whoever registered for it
            cb()

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.zero_()

    def state_dict(self):
        """The layout of ``torch.optim.Adam.state_dict()`` (one param group, parameters numbered in order), so
        checkpoints are interchangeable with the reference's optimizer (scripts/train.py)."""
        state = {}
        for i, p in enumerate(self.params):
            st = self.state.get(p)
            if st is not None:
                state[i] = {"step": torch.tensor(float(st["step"])), "exp_avg": st["exp_avg"].clone(),
                            "exp_avg_sq": st["exp_avg_sq"].clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "amsgrad": False, "maximize": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd):
        """Accepts ``FusedAdam.state_dict()`` and ``torch.optim.Adam.state_dict()`` (single param group, amsgrad off)."""
        groups = sd["param_groups"]
        if len(groups) != 1 or groups[0].get("amsgrad", False) or groups[0].get("maximize", False):
            raise ValueError("FusedAdam loads single-group Adam state without amsgrad / maximize")
        g = groups[0]
        if len(g["params"]) != len(self.params):
            raise ValueError("optimizer state has %d parameters, this optimizer %d" % (len(g["params"]), len(self.params)))
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
        self.state = {}
        for i, p in zip(g["params"], self.params):
            st = sd["state"].get(i)
            if st is None:
                continue
            mine = self._state(p)
            mine["step"] = int(st["step"].item() if torch.is_tensor(st["step"]) else st["step"])
            mine["exp_avg"].copy_(st["exp_avg"])
            mine["exp_avg_sq"].copy_(st["exp_avg_sq"])
