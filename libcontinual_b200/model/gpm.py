"""GPM (Gradient Projection Memory) on AlexNet_TRGP — host-side mirror of `core/model/gpm.py:22-206` and `core/model/backbone/alexnet.py:94-156`: same
constructor kwargs, `observe / inference / before_task / after_task / get_parameters` with the same return tuples, the same torch-RNG draw order for the
freshly initialised layers and heads, the same `torch.randperm` draw for the 125-sample selection.

Per step (`observe`, gpm.py:65-83; the Trainer calls `optimizer.zero_grad()` first and `optimizer.step()` after, trainer.py:593-606):
    AlexNet forward (dropout live in train mode) -> current task's bias-free head -> CE on labels - known -> backward -> for task > 0 every TRGP layer's
    weight gradient loses its component inside span(U_l): g <- g - g.view(out, -1) @ (U_l U_l^T)   (`lc_gpm_project_tc`: three BF16 tcgen05 GEMMs on a
    two-term split, fp32-level accuracy).
Task boundary (`after_task`, gpm.py:131-204): eval-mode pass over 125 random training samples, the representation matrix of every TRGP layer by ONE
im2col launch each (the reference fills them with a Python triple loop), SVD, threshold growth of the bases (`gpm_update_bases`, float64 like numpy's).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch
import torch.nn as nn

from .._lib import check, stream_ptr, host_acc
from ..nn_engine import ALEXNET_SPEC, AlexNetEngine
from .backbone.resnet import _Node


class AlexNet_TRGP(nn.Module):
    """Backbone factory named in the YAML recipes (`backbone.name: AlexNet_TRGP`, alexnet.py:94-122).  The device engine is created by the method class
    (which knows the head sizes); until then the module holds the initial weights, drawn exactly as the reference constructor draws them."""

    def __init__(self, dropout_rate_1: float = 0.2, dropout_rate_2: float = 0.5, **kwargs):
        super().__init__()
        if (dropout_rate_1, dropout_rate_2) != (0.2, 0.5):
            raise NotImplementedError("AlexNet_TRGP: the shipped recipes use dropout 0.2 / 0.5 (alexnet.py:96)")
        self.feat_dim = 2048
        self.max_batch = int(kwargs.get("max_batch", 128))
        self.device_arg = kwargs.get("device")
        self._init = {}
        for name, bn, cin, cout, ks, *_ in ALEXNET_SPEC:          # nn.Conv2d / nn.Linear constructors draw kaiming_uniform(a=sqrt(5)) in this order
            m = nn.Conv2d(cin, cout, ks, bias=False) if name.startswith("conv") else nn.Linear(cin, cout, bias=False)
            self._init[name + ".weight"] = m.weight.data
            self._init[bn + ".weight"] = torch.ones(cout)
            self._init[bn + ".bias"] = torch.zeros(cout)
        self.engine: Optional[AlexNetEngine] = None

    def _attach(self, engine: AlexNetEngine):
        self.engine = engine
        for name, shape in engine.layout:
            view = engine.param_view(name)
            view.copy_(self._init[name].to(view.device))
            *path, leaf = name.split(".")
            mod = self
            for part in path:
                if part not in mod._modules:
                    mod.add_module(part, _Node())
                mod = mod._modules[part]
            mod.register_parameter(leaf, nn.Parameter(view))
        self._init = None

    def _apply(self, fn, recurse=True):
        return self          # parameters are views into the engine arena: they live where the engine lives

    def forward(self, x, compute_input_matrix: bool = False):
        if self.engine is None:
            raise RuntimeError("AlexNet_TRGP runs through libcontinual_b200.model.GPM (which owns the device engine); there is no eager-PyTorch fallback")
        x = x.to(self.engine.device, torch.float32).contiguous()
        self.engine.forward(x, train=self.training, need_backward=False)
        return self.engine.features(x.shape[0]).clone()


class _HeadT(nn.Module):
    def __init__(self, view: torch.Tensor):
        super().__init__()
        self.weight = nn.Parameter(view)
        self.out_features, self.in_features = view.shape

    def _apply(self, fn, recurse=True):
        return self


class _Network(nn.Module):
    """`Network` of gpm.py:22-41: backbone + one bias-free head per task."""

    def __init__(self, backbone: AlexNet_TRGP, heads: Sequence[_HeadT]):
        super().__init__()
        self.backbone = backbone
        self.classifiers = nn.ModuleList(heads)

    def _apply(self, fn, recurse=True):
        return self


def gpm_update_bases(feature_list: List[np.ndarray], mats: Sequence[np.ndarray], task_idx: int) -> List[np.ndarray]:
    """gpm.py:170-204.  mats: float64 representation matrices [dim_l][columns].  Task 0: the leading left singular vectors that hold
    `threshold = 0.97 + 0.003 * task_idx` of the energy.  Later: remove what the stored basis already explains, and add just enough new directions
    of the residual to reach the threshold (skipped when the stored basis already does)."""
    threshold = 0.97 + task_idx * 0.003
    if task_idx == 0:
        out = []
        for act in mats:
            U, S, _ = np.linalg.svd(act, full_matrices=False)
            ratio = (S ** 2) / (S ** 2).sum()
            out.append(U[:, :int(np.sum(np.cumsum(ratio) < threshold))])
        return out
    out = list(feature_list)
    for i, act in enumerate(mats):
        _, S, _ = np.linalg.svd(act, full_matrices=False)
        total = (S ** 2).sum()
        act_hat = act - out[i] @ out[i].T @ act
        U, S, _ = np.linalg.svd(act_hat, full_matrices=False)
        ratio = (S ** 2) / total
        accumulated = (total - (S ** 2).sum()) / total
        if accumulated >= threshold:
            continue
        r = int(np.sum(np.cumsum(ratio) + accumulated < threshold)) + 1
        Ui = np.hstack((out[i], U[:, :r]))
        out[i] = Ui[:, :min(Ui.shape[0], Ui.shape[1])]
    return out


class GPM(nn.Module):
    """core/model/gpm.py:43-206."""

    def __init__(self, backbone, device, **kwargs):
        super().__init__()
        if not isinstance(backbone, AlexNet_TRGP):
            raise TypeError("libcontinual_b200.model.GPM needs the libcontinual_b200 AlexNet_TRGP backbone (gpm.py:147-150 hard-wires its geometry); "
                            "there is no eager-PyTorch fallback path")
        self.device = torch.device(device) if device is not None else None
        self.task_num, self.init_cls_num, self.inc_cls_num = kwargs["task_num"], kwargs["init_cls_num"], kwargs["inc_cls_num"]
        sizes = [self.init_cls_num] + [self.inc_cls_num] * (self.task_num - 1)
        # nn.Linear(feat_dim, n, bias=False) per task, in order (gpm.py:29-32): the same RNG draws
        head_init = [nn.Linear(backbone.feat_dim, n, bias=False).weight.data for n in sizes]
        self.engine = AlexNetEngine(max_batch=backbone.max_batch, head_sizes=sizes, device=backbone.device_arg if backbone.device_arg is not None else device)
        eng = self.engine
        backbone._attach(eng)
        heads = []
        for t, w in enumerate(head_init):
            v = eng.head_view(t)
            v.copy_(w.to(v.device))
            heads.append(_HeadT(v))
        self.network = _Network(backbone, heads)
        self.theta, self.theta_grad, self.scal = eng.theta, eng.theta_grad, eng.scal
        self._known_classes = 0
        self.cur_task = 0
        self.feature_list: List[np.ndarray] = []
        self.feature_mat: List[torch.Tensor] = []
        self._proj = None                       # (hi, lo) BF16 splits of U U^T per layer
        self._proj_scratch = {}
        self.layers = [n for n, *_ in ALEXNET_SPEC]             # 3 conv, then 2 linear (gpm.py:58-61)
        self._bn_frozen = False

    # ---- parameters / optimizer glue ---------------------------------------------------------------------------------------------------------------
    def get_parameters(self, config):
        return self.network.parameters()

    def _trainable(self):
        """(parameter, arena offset) of everything the current task updates: all backbone weights, BN affine until task 0 ends (gpm.py:126-129),
        the current head (the other heads receive no gradient: gpm.py:70 uses logits[cur_task] only)."""
        eng = self.engine
        out = []
        for name, _ in eng.layout:
            if self._bn_frozen and name.startswith("bn"):
                continue
            mod = self.network.backbone
            for part in name.split("."):
                mod = getattr(mod, part)
            out.append((mod, eng.param_off[name][0]))
        out.append((self.network.classifiers[self.cur_task].weight, eng.head_off[self.cur_task]))
        return out

    def active_ranges(self):
        """Contiguous arena ranges the flat optimizer updates (`optim.FlatSGD`)."""
        spans = sorted((off, off + p.numel()) for p, off in self._trainable())
        merged = []
        for lo, hi in spans:
            if merged and merged[-1][1] == lo:
                merged[-1][1] = hi
            else:
                merged.append([lo, hi])
        return [(lo, hi) for lo, hi in merged]

    # ---- task boundaries ----------------------------------------------------------------------------------------------------------------------------
    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.cur_task = task_idx
        if task_idx == 1:
            self._known_classes += self.init_cls_num
        elif task_idx > 1:
            self._known_classes += self.inc_cls_num
        if task_idx > 0:
            eng = self.engine
            self.feature_mat = [torch.tensor(f @ f.T, dtype=torch.float32, device=eng.device) for f in self.feature_list]        # gpm.py:124
            hi_lo = []
            for Mx in self.feature_mat:
                Ms = ((Mx + Mx.T) * 0.5).contiguous()
                hi = torch.empty(Ms.shape, device=eng.device, dtype=torch.bfloat16); lo = torch.empty_like(hi)
                check(eng.lib.lc_split_bf16(Ms.data_ptr(), hi.data_ptr(), lo.data_ptr(), Ms.numel(), stream_ptr()), "lc_split_bf16")
                hi_lo.append((hi, lo))
            self._proj = hi_lo
            for name, p in self.network.named_parameters():          # gpm.py:126-129
                p.requires_grad_(True)
                if "bn" in name:
                    p.requires_grad_(False)
            self._bn_frozen = True

    @torch.no_grad()
    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        eng = self.engine
        x = torch.cat([batch["image"].to(eng.device, torch.float32) for batch in train_loader], dim=0)
        sel = torch.randperm(x.size(0))[:125]                          # the reference's CPU draw (gpm.py:139-142)
        x = x[sel.to(x.device)].contiguous()
        was = self.network.training
        self.network.eval()
        mats = eng.representation_matrices(x)
        self.network.train(was)
        self.feature_list = gpm_update_bases(self.feature_list, [m.double().cpu().numpy() for m in mats], task_idx)

    # ---- the step -------------------------------------------------------------------------------------------------------------------------------------
    def _to_device(self, data):
        x = data["image"].to(self.engine.device, torch.float32, non_blocking=True).contiguous()
        y = data["label"].to(self.engine.device, torch.int64, non_blocking=True).contiguous()
        return x, y

    def _launch_step(self, x, y, clip: bool = True):
        """Everything `observe` puts on the stream (no host synchronisation): capturable into a CUDA graph."""
        eng = self.engine
        B = x.shape[0]
        train = self.network.training
        eng._train_step = train
        if train:
            check(eng.lib.lc_nn_rng_advance(eng.rng.data_ptr(), stream_ptr()), "lc_nn_rng_advance")
        eng.forward(x, train=train, need_backward=True)
        eng.heads_forward(B, self.cur_task)
        torch.sub(y, self._known_classes, out=eng.yshift[:B])
        eng.loss_backward_head(eng.yshift, B, self.cur_task)
        eng.backward(B)
        if self.cur_task > 0:
            for i, name in enumerate(self.layers):                     # gpm.py:76-81
                self._project(i, eng.param_view(name + ".weight", eng.theta_grad))
        eng.launches += 2

    def _project(self, i: int, grad: torch.Tensor):
        eng = self.engine
        rows, dim = grad.shape[0], grad.numel() // grad.shape[0]
        if (rows, dim) not in self._proj_scratch:
            self._proj_scratch[(rows, dim)] = (torch.empty(rows, dim, device=eng.device, dtype=torch.bfloat16),
                                               torch.empty(rows, dim, device=eng.device, dtype=torch.bfloat16))
        gh, gl = self._proj_scratch[(rows, dim)]
        hi, lo = self._proj[i]
        check(eng.lib.lc_gpm_project_tc(grad.data_ptr(), hi.data_ptr(), lo.data_ptr(), rows, dim, gh.data_ptr(), gl.data_ptr(), eng.err.data_ptr(), stream_ptr()),
              "lc_gpm_project_tc")
        eng.launches += 5

    def train(self, mode: bool = True):
        super().train(mode)
        self.network.training = mode
        self.network.backbone.training = mode
        return self

    def observe(self, data):
        x, y = self._to_device(data)
        self._launch_step(x, y)
        eng = self.engine
        for p, off in self._trainable():                               # loss.backward() happened inside (trainer.py:593-596): grads are in place
            p.grad = eng.theta_grad[off:off + p.numel()].view(p.shape)
        B = x.shape[0]
        acc = float(eng.scal[1].item()) / B
        return eng.pred[:B].clone(), acc, eng.scal[0]

    @torch.no_grad()
    def inference(self, data, task_id: int = -1):
        """gpm.py:85-111: task-aware (one head, prediction shifted by the head's class offset) or task-agnostic (all heads side by side)."""
        x, y = self._to_device(data)
        eng = self.engine
        B = x.shape[0]
        eng.forward(x, train=self.network.training, need_backward=False)
        if task_id > -1:
            bias = 0 if task_id == 0 else self.init_cls_num + (task_id - 1) * self.inc_cls_num
            preds = eng.heads_forward(B, task_id).max(1)[1] + bias
        else:
            preds = eng.heads_forward(B, None).max(1)[1]
        return preds, host_acc(self, preds.eq(y).sum(), y.size(0))
