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
                    if not (0 <= off < eng.n_total):
                        continue          # a parameter outside the arena (never part of the fused step, e.g. the reference's unused backbone.fc)
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


class Adam(torch.optim.Optimizer):
    """torch.optim.Adam (config/l2p-vit-cifar100-b10-10-10.yaml:37-42) as one kernel over a model's flat trainable arena
    (`model.theta` / `model.theta_grad`, e.g. `libcontinual_b200.model.l2p.L2P`).  State (exp_avg, exp_avg_sq, step) is rebuilt
    with the optimizer every task, like the reference (trainer.py:294)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, *, model=None):
        if model is None or not hasattr(model, "theta"):
            raise ValueError("libcontinual_b200.optim.Adam needs model= with a flat .theta / .theta_grad arena")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self.model = model
        self.m = torch.zeros_like(model.theta)
        self.v = torch.zeros_like(model.theta)
        self.hp = torch.zeros(12, device=model.theta.device)      # [0..7] fp32 (see `hyper`), [8..11] = (beta1, beta2) as float64 for `lc_adam_tick`
        self.t = 0
        from . import _lib
        self.lib = _lib.load()

    def hyper(self, t: int):
        g = self.param_groups[0]
        b1, b2 = g["betas"]
        return [float(g["lr"]), float(b1), float(b2), float(g["eps"]), float(g["weight_decay"]), 1.0 - b1 ** t, 1.0 - b2 ** t, float(t)]

    def zero_grad(self, set_to_none: bool = True):
        # gradients are overwritten (not accumulated) by the model's backward kernels
        for grp in self.param_groups:
            for p in grp["params"]:
                p.grad = None

    @torch.no_grad()
    def step(self, closure=None):
        mdl = self.model
        self.t += 1
        self.hp[:8].copy_(torch.tensor(self.hyper(self.t), dtype=torch.float32), non_blocking=False)
        base = mdl.theta.data_ptr()
        for grp in self.param_groups:
            for p in grp["params"]:
                if p.grad is None:
                    raise RuntimeError("Adam.step() before observe(): no gradient")
                off = (p.data_ptr() - base) // 4
                if p.grad.data_ptr() != mdl.theta_grad.data_ptr() + off * 4:         # foreign gradient: gather it into the arena layout
                    mdl.theta_grad[off:off + p.numel()].copy_(p.grad.reshape(-1))
        self.launch()
        return None

    def launch(self, tick: bool = False):
        """The update kernel alone (hp already on the device): what a captured step replays.  tick=True puts the device-side step counter
        (`lc_adam_tick`: t += 1 and the two bias corrections) in front of it, so that a replayed graph needs no per-step host staging."""
        mdl = self.model
        if tick:
            check(self.lib.lc_adam_tick(self.hp.data_ptr(), torch.cuda.current_stream().cuda_stream), "adam_tick")
        check(self.lib.lc_adam(mdl.theta.data_ptr(), mdl.theta_grad.data_ptr(), self.m.data_ptr(), self.v.data_ptr(), mdl.theta.numel(),
                               self.hp.data_ptr(), torch.cuda.current_stream().cuda_stream), "adam")


class FlatSGD(torch.optim.Optimizer):
    """torch.optim.SGD (momentum, weight decay; config/InfLoRA_opt-vit-imagenetr-b20-20-10.yaml:44-48) over the ACTIVE ranges of a model's flat
    trainable arena (`model.theta`, `model.theta_grad`, `model.active_ranges()`), one fused kernel per range.  Momentum buffers are rebuilt
    with the optimizer every task, like the reference (trainer.py:294)."""

    def __init__(self, params, lr=1e-3, momentum=0.0, weight_decay=0.0, *, model=None):
        if model is None or not hasattr(model, "theta") or not hasattr(model, "active_ranges"):
            raise ValueError("libcontinual_b200.optim.FlatSGD needs model= with .theta / .theta_grad / .active_ranges()")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay))
        self.model = model
        self.buf = torch.zeros_like(model.theta)
        self.hp = torch.zeros(4, device=model.theta.device)
        self._hp_host = None
        from . import _lib
        self.lib = _lib.load()

    def sync_hp(self):
        g = self.param_groups[0]
        hp = (float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]))
        if hp != self._hp_host:
            self.hp[:3] = torch.tensor(hp, device=self.hp.device)
            self._hp_host = hp

    @torch.no_grad()
    def step(self, closure=None):
        mdl = self.model
        self.sync_hp()
        ag = getattr(mdl, "autograd_grads", None)
        base = mdl.theta.data_ptr()
        src = None
        for grp in self.param_groups:
            for p in grp["params"]:
                if p.grad is None:
                    continue
                off = (p.data_ptr() - base) // 4
                if ag is not None and p.grad.data_ptr() == ag.data_ptr() + off * 4:
                    src = ag                                                   # the copy loss.backward() produced, already in arena layout
                elif p.grad.data_ptr() != mdl.theta_grad.data_ptr() + off * 4:
                    mdl.theta_grad[off:off + p.numel()].copy_(p.grad.reshape(-1))     # foreign gradient: gather it
        self.launch(mdl.theta_grad if src is None else src)
        return None

    def launch(self, grads=None):
        """The update kernels alone (hp already on the device): what a captured step replays."""
        mdl = self.model
        g = mdl.theta_grad if grads is None else grads
        st = torch.cuda.current_stream().cuda_stream
        for lo, hi in mdl.active_ranges():          # a range may start off a 16-byte boundary (10-class head bias): lc_sgd_momentum falls back to its scalar kernel
            check(self.lib.lc_sgd_momentum(mdl.theta.data_ptr() + 4 * lo, g.data_ptr() + 4 * lo, self.buf.data_ptr() + 4 * lo, hi - lo, self.hp.data_ptr(), st),
                  "sgd_momentum")
