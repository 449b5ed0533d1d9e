"""Flat-arena optimizers: one fused kernel over the whole parameter arena instead of torch's per-tensor foreach kernels
(reference: torch.optim.SGD / Adam built by trainer.py:159-182 and rebuilt every task at :294).

They subclass torch.optim.Optimizer so that LR schedulers (`param_groups[i]['lr']`) keep working."""
from __future__ import annotations

import torch

from ._lib import check


class SGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.1, momentum=0.0, weight_decay=0.0, *, engine=None, model=None):
        if engine is None:
            engine = getattr(model, "engine", None)
        if engine is None:
            raise ValueError("libcontinual_b200.optim.SGD needs engine= (or model= with an .engine)")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.engine = engine
        self.buf = torch.zeros_like(engine.params)
        self.hp = torch.zeros(4, device=engine.device)
        self._hp_host = None
        self._check_views = True

    @torch.no_grad()
    def step(self, closure=None):
        g = self.param_groups[0]
        hp = (float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]))
        if hp != self._hp_host:
            self.hp[:3] = torch.tensor(hp, device=self.hp.device)
            self._hp_host = hp
        eng = self.engine
        ag = eng.autograd_grads
        p0 = g["params"][0]
        if ag is not None and p0.grad is not None and p0.grad.data_ptr() == ag.data_ptr():
            src = ag          # fast path: p.grad are the views of the flat copy that backward() produced (clipping included)
        else:
            # generic path: gather whatever autograd left in p.grad back into the arena layout
            src = eng.grads
            base = eng.params.data_ptr()
            for grp in self.param_groups:
                for p in grp["params"]:
                    off = (p.data_ptr() - base) // 4
                    if p.grad is None:
                        src[off:off + p.numel()].zero_()
                    else:
                        src[off:off + p.numel()].copy_(p.grad.reshape(-1))
        eng.sgd_step(self.buf, self.hp, grads=src, frozen=self._frozen_range())
        return None

    def _frozen_range(self):
        """A param group with lr == 0 and weight_decay == 0 (LUCIR's old-class embedding, lucir.py:229-240) maps to one contiguous
        arena range that the fused kernel skips."""
        lo = hi = None
        base = self.engine.params.data_ptr()
        for grp in self.param_groups[1:]:
            if float(grp["lr"]) == 0.0 and float(grp["weight_decay"]) == 0.0:
                for p in grp["params"]:
                    off = (p.data_ptr() - base) // 4
                    lo = off if lo is None else min(lo, off)
                    hi = off + p.numel() if hi is None else max(hi, off + p.numel())
            else:
                raise ValueError("libcontinual_b200.optim.SGD: extra param groups must be frozen (lr = 0, weight_decay = 0)")
        return None if lo is None else (int(lo), int(hi))
