"""InfLoRA_OPT on ViT-B/16 — mirror of the reference plugin surface (core/model/InfLoRA_opt.py:46-459: `SiNet`, `InfLoRA_OPT`;
core/model/backbone/transformer.py:199-274 `MultiHeadAttention_LoRA`) on top of `ViTEngine`.

    backbone = vit_pt_imnet(pretrained=False, state=<VisionTransformer state_dict>, attn_layer="MultiHeadAttention_LoRA", lora_rank=10)
    model    = InfLoRA_OPT(backbone, device, init_cls_num=20, inc_cls_num=20, task_num=10, lame=1.0, lamb=0.95, embd_dim=768,
                           use_ca=False, dataset="imagenet-r")
    model.before_task(t, buffer, train_loader, test_loaders)        # input-matrix pass + SVD -> lora_A; lora_B <- 0
    pred, acc, loss = model.observe(batch); optimizer.zero_grad(); loss.backward(); optimizer.step()      # trainer.py:601-606
    model.after_task(t, buffer, train_loader, test_loaders)         # merge_weight + DualGPM feature update

What is different from the reference underneath: the adapters are merged into the BF16 GEMM operands once per step (one launch for all
12 blocks), the backbone backward carries only token gradients (frozen weights) and the adapter gradients are formed in rank-10 form from
the saved down-projections (`lc_lora_bgrad_rows`) instead of dense 768 x 768 weight gradients.  Trainables (lora_B_k / lora_B_v of every
block + all task heads) live in one flat arena `theta`; `libcontinual_b200.optim.FlatSGD` updates the active ranges.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from .._lib import LcError, check, stream_ptr, host_acc
from ..vit_engine import DIM, LoraState
from .l2p import ViTZoo


class _FlatLoss(torch.autograd.Function):
    """Scalar loss whose backward hands autograd the gradients that the fused step already left in the owner's flat arena, so the
    reference order observe -> zero_grad -> loss.backward() -> optimizer.step() (trainer.py:601-606) runs unchanged."""

    @staticmethod
    def forward(ctx, owner, loss_value, *params):
        ctx.owner = owner
        return loss_value.clone()

    @staticmethod
    def backward(ctx, g):
        owner = ctx.owner
        owner.autograd_grads = owner.theta_grad * g       # a copy: p.grad must not alias the arena the next step overwrites
        return (None, None, *[owner._grad_view(p, owner.autograd_grads) for p in owner.trainable_params()])


class SiNet(nn.Module):
    """InfLoRA_opt.py:46-129: backbone + one nn.Linear head per task (`classifier_pool`).  The heads are views of the owner's flat arena."""

    def __init__(self, backbone: ViTZoo, device, **kwargs):
        super().__init__()
        self._cur_task_id = -1
        self.backbone = backbone
        self.device = device
        seed = int(os.environ.get("PYTHONHASHSEED", "0") or 0)
        torch.manual_seed(seed)                                              # `_set_random(os.environ["PYTHONHASHSEED"])`, InfLoRA_opt.py:55
        sizes = [kwargs["init_cls_num"]] + [kwargs["inc_cls_num"]] * (kwargs["task_num"] - 1)
        self.classifier_pool = nn.ModuleList([nn.Linear(kwargs["embd_dim"], c, bias=True) for c in sizes])

    def update_fc(self, train_loader):
        self._cur_task_id += 1


class InfLoRA_OPT(nn.Module):
    def __init__(self, backbone: ViTZoo, device, **kwargs):
        super().__init__()
        if not isinstance(backbone, ViTZoo):
            raise LcError("InfLoRA_OPT needs a libcontinual_b200 ViTZoo backbone (the CLIP branch of the reference is not on the CUDA path)")
        self.device = torch.device(device)
        self.init_cls_num, self.inc_cls_num, self.task_num = kwargs["init_cls_num"], kwargs["inc_cls_num"], kwargs["task_num"]
        self.lame, self.lamb = kwargs["lame"], kwargs["lamb"]
        self._known_classes = 0
        self.feature_list: List[np.ndarray] = []
        self.project_type: List[str] = []
        self._dataset = kwargs.get("dataset", "imagenet-r")
        self._use_class_alignment = kwargs.get("use_ca", False)
        if self._use_class_alignment:
            raise NotImplementedError("use_ca=True (classifier alignment) is outside the per-step hot path and not built")
        self.engine = eng = backbone.engine
        self.rank = int(getattr(backbone, "lora_rank", kwargs.get("lora_rank", 10)))
        self._network = SiNet(backbone, device, **kwargs)
        L, r, dev = eng.depth, self.rank, eng.dev
        sizes = [self.init_cls_num] + [self.inc_cls_num] * (self.task_num - 1)
        self.total_cls = sum(sizes)
        self.cls_lo = [sum(sizes[:t]) for t in range(self.task_num)]
        self.cls_n = sizes
        # flat arena: [lora_B (L, 2, 768, r) | head weights (total_cls, 768) | head biases (total_cls)]
        self.nB = L * 2 * DIM * r
        self.oW, self.ob = self.nB, self.nB + self.total_cls * DIM
        self.theta = torch.zeros(self.ob + self.total_cls, device=dev)
        self.theta_grad = torch.zeros_like(self.theta)
        self.lora_B = self.theta[:self.nB].view(L, 2, DIM, r)
        self.heads_W = self.theta[self.oW:self.ob].view(self.total_cls, DIM)
        self.heads_b = self.theta[self.ob:]
        eng.lora = LoraState(eng, (1, 2), r, self.lora_B, self.theta_grad[:self.nB].view(L, 2, DIM, r))
        # nn.Parameter views (names follow the reference modules: lora_B_k / lora_B_v per block, classifier_pool.{t}.weight / .bias)
        self.lora_B_k = nn.ParameterList([nn.Parameter(self.lora_B[i, 0], requires_grad=False) for i in range(L)])
        self.lora_B_v = nn.ParameterList([nn.Parameter(self.lora_B[i, 1], requires_grad=False) for i in range(L)])
        for t, head in enumerate(self._network.classifier_pool):
            lo, n = self.cls_lo[t], self.cls_n[t]
            self.heads_W[lo:lo + n].copy_(head.weight.detach()); self.heads_b[lo:lo + n].copy_(head.bias.detach())
            head.weight.data = self.heads_W[lo:lo + n]; head.bias.data = self.heads_b[lo:lo + n]
            head.weight.requires_grad_(False); head.bias.requires_grad_(False)
        self.cur_task = -1
        self._bufs = {}
        self.scal = torch.zeros(8, device=dev)
        self.autograd_grads: Optional[torch.Tensor] = None

    # ---- flat-arena bookkeeping --------------------------------------------------------------------
    def _grad_view(self, p: torch.Tensor, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        off = (p.data_ptr() - self.theta.data_ptr()) // 4
        return (self.theta_grad if arena is None else arena)[off:off + p.numel()].view(p.shape)

    def trainable_params(self) -> List[nn.Parameter]:
        head = self._network.classifier_pool[self.cur_task]
        return list(self.lora_B_k) + list(self.lora_B_v) + [head.weight, head.bias]

    def active_ranges(self):
        """Element ranges of `theta` that train in the current task: all lora_B, the current head's weight rows and bias."""
        lo, n = self.cls_lo[self.cur_task], self.cls_n[self.cur_task]
        return [(0, self.nB), (self.oW + lo * DIM, self.oW + (lo + n) * DIM), (self.ob + lo, self.ob + lo + n)]

    def get_parameters(self, config):
        """InfLoRA_opt.py:458-459 returns every parameter; torch.optim skips those without a gradient.  Only the trainable ones are exposed."""
        return self.trainable_params()

    def _batch_bufs(self, B):
        if B not in self._bufs:
            dev = self.engine.dev
            self._bufs[B] = dict(logits=torch.zeros(B, self.total_cls, device=dev), dlogits=torch.zeros(B, self.total_cls, device=dev),
                                 pred=torch.zeros(B, dtype=torch.int64, device=dev), dfeat=torch.zeros(B, DIM, device=dev))
        return self._bufs[B]

    def _to_device(self, data):
        x = data["image"].to(self.engine.dev, torch.float32, non_blocking=True).contiguous()
        y = data["label"].to(self.engine.dev, torch.int64, non_blocking=True).contiguous()
        return x, y

    # ---- the step ------------------------------------------------------------------------------------
    def _launch_step(self, x, y, clip: bool = True):
        """Everything observe() puts on the stream (capturable): adapter merge, forward, head, CE over the task's logits, backward to
        every lora_B and the head.  y holds absolute labels; the task head covers [known, known + n)."""
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        bb = self._batch_bufs(B)
        lo, n = self.cls_lo[self.cur_task], self.cls_n[self.cur_task]
        eng.lora_merge()
        ws = eng.forward(x, None, save=True)
        feat = eng.pooled(ws, 0)
        C = self.total_cls
        check(lib.lc_linear_head(feat.data_ptr(), self.heads_W[lo].data_ptr(), self.heads_b[lo:].data_ptr(), B, n, DIM, bb["logits"][:, lo:].data_ptr(), C, st),
              "linear_head")
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, y.data_ptr(), B, lo, lo + n, None, 0.0, bb["dlogits"].data_ptr(), bb["pred"].data_ptr(),
                                    self.scal.data_ptr(), st), "loss_ce_masked")
        gW = self.theta_grad[self.oW + lo * DIM:]
        gb = self.theta_grad[self.ob + lo:]
        check(lib.lc_linear_head_backward(bb["dlogits"][:, lo:].data_ptr(), C, feat.data_ptr(), self.heads_W[lo].data_ptr(), n, B, DIM, gW.data_ptr(),
                                          gb.data_ptr(), bb["dfeat"].data_ptr(), st), "linear_head_backward")
        eng.backward_tokens(ws, bb["dfeat"], 0, to_tokens=False)
        eng.launches += 3
        return bb

    def observe(self, data):
        x, y = self._to_device(data)
        bb = self._launch_step(x, y)
        B = x.shape[0]
        acc = float(self.scal[1].item()) / B
        return bb["pred"] - self._known_classes, acc, _FlatLoss.apply(self, self.scal[0], *self.trainable_params())

    @torch.no_grad()
    def inference(self, data):
        """InfLoRA_opt.py:191-205: logits of every head seen so far, argmax."""
        x, y = self._to_device(data)
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        bb = self._batch_bufs(B)
        if eng.lora.active:
            eng.lora_merge()
        ws = eng.forward(x, None, save=False)
        feat = eng.pooled(ws, 0)
        seen = self.cls_lo[self.cur_task] + self.cls_n[self.cur_task]
        C = self.total_cls
        check(lib.lc_linear_head(feat.data_ptr(), self.heads_W.data_ptr(), self.heads_b.data_ptr(), B, seen, DIM, bb["logits"].data_ptr(), C, st), "linear_head")
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, y.data_ptr(), B, 0, seen, None, 0.0, bb["dlogits"].data_ptr(), bb["pred"].data_ptr(),
                                    self.scal.data_ptr(), st), "argmax")
        eng.launches += 2
        return bb["pred"], host_acc(self, self.scal[1], B)

    # ---- task boundaries -------------------------------------------------------------------------------
    @torch.no_grad()
    def input_matrices(self, loader) -> torch.Tensor:
        """`update_input_matrix` over a loader (InfLoRA_opt.py:243-245): per block the token mean of h h^T, [L, 768, 768] on the device."""
        eng = self.engine
        if eng.lora.active:
            eng.lora_merge()
        eng.input_matrix_begin()
        for batch in loader:
            x = (batch["image"] if isinstance(batch, dict) else batch).to(eng.dev, torch.float32).contiguous()
            eng.forward(x, None, save=False)
        return eng.input_matrix_end()

    def start_task(self, task_idx: int, A: Optional[torch.Tensor] = None):
        """before_task without the loader pass: bookkeeping, lora_B <- 0, requires_grad flags; `A` [L, 2, r, 768] installs given bases."""
        if task_idx == 1:
            self._known_classes = self.init_cls_num
        elif task_idx > 1:
            self._known_classes += self.inc_cls_num
        self._network.update_fc(None)
        self.cur_task = task_idx
        self.lora_B.zero_()                                                   # init_param(): zeros_(lora_B), transformer.py:230-231
        self.engine.lora.active = True
        for t, head in enumerate(self._network.classifier_pool):
            head.weight.requires_grad_(t == task_idx); head.bias.requires_grad_(t == task_idx)
        for p in list(self.lora_B_k) + list(self.lora_B_v):
            p.requires_grad_(True)
        if A is not None:
            self.engine.lora.set_A(A.to(self.engine.dev))

    @torch.no_grad()
    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        """InfLoRA_opt.py:207-264: the new adapter's down-projection is the top-r left singular basis of the task's input matrix, taken
        inside (project_type 'retain') or outside ('remove') the subspace spanned by the previous tasks' features."""
        self.start_task(task_idx)
        # lora_A enters the input-matrix pass only through B A with B = 0, so the pass sees the merged weights of the previous tasks
        cur = self.input_matrices(train_loader)
        L, r = self.engine.depth, self.rank
        A = torch.empty(L, 2, r, DIM, device=self.engine.dev)
        for i in range(L):
            m = cur[i]
            if task_idx > 0:
                F_i = torch.from_numpy(np.ascontiguousarray(self.feature_list[i])).to(m)
                inside = F_i @ (F_i.T @ m)
                m = m - inside if self.project_type[i] == "remove" else inside
            U, _, _ = torch.linalg.svd(m, full_matrices=False)
            A[i, 0] = A[i, 1] = U[:, :r].T / math.sqrt(3)
        self.engine.lora.set_A(A)

    @torch.no_grad()
    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        """InfLoRA_opt.py:266-276: fold the adapters into the weights (`merge_weight`), then update the DualGPM feature subspaces."""
        eng = self.engine
        eng.lora_merge(w_out=True)
        eng.lora.active = False
        self._update_feature(task_idx, train_loader)

    @torch.no_grad()
    def _update_feature(self, task_idx, train_loader):
        acts = self.input_matrices(train_loader).cpu().numpy()
        dualgpm_update(acts, self.feature_list, self.project_type, task_idx, self.task_num, self.lame, self.lamb)


def dualgpm_update(acts, feature_list: List[np.ndarray], project_type: List[str], task_idx: int, task_num: int, lame: float, lamb: float):
    """DualGPM bookkeeping of InfLoRA_opt.py:278-362 (host side, once per task; in place on the two lists): per block keep either a basis of the
    input subspace used so far ('remove') or of its complement ('retain'), grown / shrunk so that the captured input-matrix energy crosses
    `threshold`; a 'remove' basis that outgrows half the dimension is swapped for its orthogonal complement.  `acts`: [L, D, D] input matrices."""
    threshold = (lame - lamb) * task_idx / task_num + lamb
    for i in range(acts.shape[0]):
        act = acts[i]
        if task_idx == 0:
            U, S, _ = np.linalg.svd(act, full_matrices=False)
            ratio = S ** 2 / (S ** 2).sum()
            r = max(int(np.sum(np.cumsum(ratio) < threshold)), 1)
            assert r < act.shape[0] / 2
            feature_list.append(U[:, :r])
            project_type.append("remove")
            continue
        total = (np.linalg.svd(act, compute_uv=False) ** 2).sum()
        F_i = feature_list[i]
        inside = (F_i @ F_i.T).astype(np.float32) @ act
        if project_type[i] == "remove":
            U, S, _ = np.linalg.svd(act - inside, full_matrices=False)
            ratio = S ** 2 / total
            kept = (total - (S ** 2).sum()) / total
            if kept < threshold:
                r = int(np.sum(np.cumsum(ratio) + kept < threshold)) + 1
                grown = np.hstack((F_i, U[:, :r]))
                feature_list[i] = grown[:, :min(grown.shape)]
        else:
            U, S, _ = np.linalg.svd(inside, full_matrices=False)
            ratio = S ** 2 / total
            kept = (S ** 2).sum() / total
            if kept >= 1 - threshold:
                r = int(np.sum(kept - np.cumsum(ratio) >= 1 - threshold)) + 1
                shrunk = F_i - U[:, :r] @ U[:, :r].T @ F_i
                U2, _, _ = np.linalg.svd(shrunk)
                feature_list[i] = U2[:, :F_i.shape[1] - r]
    for i, F_i in enumerate(feature_list):
        if project_type[i] == "remove" and F_i.shape[1] > F_i.shape[0] / 2:
            U, _, _ = np.linalg.svd(F_i)
            feature_list[i] = U[:, F_i.shape[1]:]
            project_type[i] = "retain"
        elif project_type[i] == "retain":
            assert F_i.shape[1] <= F_i.shape[0] / 2
