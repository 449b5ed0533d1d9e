"""InfLoRA (original formulation) — mirror of the reference plugin surface (core/model/InfLoRA.py:36-330 `InfLoRA`; core/model/backbone/SiNet.py:60-140
`SiNet_vit`; core/model/backbone/vit_inflora.py:176-252 `Attention_LoRA`) on top of `ViTEngine`.

    backbone = SiNet_vit(total_sessions=10, rank=10, init_cls=10, embd_dim=768, state=<timm ViT-B/16 state_dict>, device=dev)
    model    = InfLoRA(backbone, 768, 100, inc_cls_num=10, device=dev, lame=1.0, lamb=0.6, total_sessions=10)

Differences from InfLoRA_OPT that matter to the kernels: the adapters of ALL tasks so far stay separate and their sum enters every forward
(`weight_k = sum_t B_t A_t`, vit_inflora.py:236-240) instead of being merged into the weights after each task; the backbone is the timm-style ViT
(LayerNorm eps 1e-6 in the blocks, state-dict keys `blocks.{i}.norm1/attn/norm2/mlp`); every head has `init_cls` outputs.  Here the adapters are
stacked along the rank axis and folded into the BF16 GEMM operands once per step (`lc_lora_merge`, R = rank * (t + 1)); the current task's lora_B
gradient is formed in rank form as in InfLoRA_OPT.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .._lib import LcError, check, stream_ptr, host_acc
from ..vit_engine import DIM, StackedLoraState
from .inflora import _FlatLoss
from .l2p import ViTZoo


class SiNet_vit(nn.Module):
    """SiNet.py:60-140: image encoder (frozen ViT-B/16 with per-task k / v adapters) + one `init_cls`-way nn.Linear head per session."""

    def __init__(self, state=None, device="cuda:0", **args):
        super().__init__()
        self.total_sessions, self.rank, self.class_num, self.embd_dim = args["total_sessions"], args["rank"], args["init_cls"], args["embd_dim"]
        self.image_encoder = ViTZoo(pretrained=state is None and args.get("pretrained", False), state=state, device=device)
        self.image_encoder.engine.block_ln_eps = 1e-6            # timm-style blocks (vit_inflora.py: norm_layer = LayerNorm(eps=1e-6))
        self.engine = self.image_encoder.engine
        self.classifier_pool = nn.ModuleList([nn.Linear(self.embd_dim, self.class_num, bias=True) for _ in range(self.total_sessions)])
        self.classifier_pool_backup = nn.ModuleList([nn.Linear(self.embd_dim, self.class_num, bias=True) for _ in range(self.total_sessions)])
        self.numtask = 0

    def update_fc(self, nb_classes):
        self.numtask += 1


class InfLoRA(nn.Module):
    def __init__(self, backbone: SiNet_vit, feat_dim, num_class, **kwargs):
        super().__init__()
        if not isinstance(backbone, SiNet_vit):
            raise LcError("InfLoRA needs a libcontinual_b200 SiNet_vit backbone")
        self._network = backbone
        self.engine = eng = backbone.engine
        self.num_class = num_class
        self._total_classes = self._known_classes = 0
        self._cur_task = -1
        self.inc_cls_num = kwargs["inc_cls_num"]
        self.device = torch.device(kwargs["device"])
        self.feature_list: List[np.ndarray] = []
        self.project_type: List[str] = []
        self.feature_mat: List[torch.Tensor] = []
        self.lame, self.lamb, self.total_sessions = kwargs["lame"], kwargs["lamb"], kwargs["total_sessions"]
        L, r, dev = eng.depth, backbone.rank, eng.dev
        self.rank = r
        S, c = backbone.total_sessions, backbone.class_num
        self.cp = (c + 3) // 4 * 4                    # bias slots are padded to 16 bytes so that every head's slice is float4-aligned for the fused SGD
        self.total_heads = S * c
        self.nB = L * 2 * DIM * r
        self.oW, self.ob = self.nB, self.nB + self.total_heads * DIM
        self.theta = torch.zeros(self.ob + S * self.cp, device=dev)
        self.theta_grad = torch.zeros_like(self.theta)
        self.B_cur = self.theta[:self.nB].view(L, 2, DIM, r)
        self.heads_W = self.theta[self.oW:self.ob].view(self.total_heads, DIM)
        self.heads_b = self.theta[self.ob:].view(S, self.cp)
        for t, head in enumerate(backbone.classifier_pool):
            self.heads_W[t * c:(t + 1) * c].copy_(head.weight.detach()); self.heads_b[t, :c].copy_(head.bias.detach())
            head.weight.data = self.heads_W[t * c:(t + 1) * c]; head.bias.data = self.heads_b[t, :c]
            head.weight.requires_grad_(False); head.bias.requires_grad_(False)
        self.A_old: Optional[torch.Tensor] = None
        self.B_old: Optional[torch.Tensor] = None
        self.A_cur = torch.zeros(L, 2, r, DIM, device=dev)
        self.params: List[nn.Parameter] = []
        self._bufs = {}
        self.scal = torch.zeros(8, device=dev)
        self.autograd_grads: Optional[torch.Tensor] = None

    # ---- flat-arena bookkeeping --------------------------------------------------------------------
    def _grad_view(self, p, arena=None):
        off = (p.data_ptr() - self.theta.data_ptr()) // 4
        return (self.theta_grad if arena is None else arena)[off:off + p.numel()].view(p.shape)

    def trainable_params(self):
        return self.params

    def get_parameters(self, config):
        return self.params

    def active_ranges(self):
        c, t = self._network.class_num, self._cur_task
        return [(0, self.nB), (self.oW + t * c * DIM, self.oW + (t + 1) * c * DIM), (self.ob + t * self.cp, self.ob + t * self.cp + c)]

    def _batch_bufs(self, B):
        if B not in self._bufs:
            dev = self.engine.dev
            self._bufs[B] = dict(logits=torch.zeros(B, self.total_heads, device=dev), dlogits=torch.zeros(B, self.total_heads, device=dev),
                                 pred=torch.zeros(B, dtype=torch.int64, device=dev), dfeat=torch.zeros(B, DIM, device=dev),
                                 yrel=torch.zeros(B, dtype=torch.int64, device=dev))
        return self._bufs[B]

    def _images(self, x):
        x = x.to(self.engine.dev, torch.float32, non_blocking=True)
        if x.shape[-1] != 224:                                   # InfLoRA.py:151,199: inputs are bilinearly resized to 224 before the encoder (input pipeline)
            x = F.interpolate(x, size=224, mode="bilinear", align_corners=False)
        return x.contiguous()

    # ---- task boundaries -------------------------------------------------------------------------------
    def start_task(self, A: Optional[torch.Tensor] = None):
        """before_task without the loader pass (InfLoRA.py:104-138): counters, the previous adapter joins the frozen stack, lora_B of the new one is
        zero (`init_param`, vit_inflora.py:203-208), requires_grad flags."""
        eng, L = self.engine, self.engine.depth
        if self._cur_task >= 0:
            A_prev, B_prev = self.A_cur.clone(), self.B_cur.clone()
            self.A_old = A_prev if self.A_old is None else torch.cat([self.A_old, A_prev], dim=2)
            self.B_old = B_prev if self.B_old is None else torch.cat([self.B_old, B_prev], dim=3)
        self._known_classes = self._total_classes
        self._cur_task += 1
        self._total_classes = self._known_classes + self.inc_cls_num
        self._network.update_fc(self._total_classes)
        self.B_cur.zero_()
        t = self._cur_task
        eng.lora = StackedLoraState(eng, self.rank, t + 1, self.B_cur, self._grad_view(self.B_cur), self.A_old, self.B_old)
        eng.lora.active = True
        for i, head in enumerate(self._network.classifier_pool):
            head.weight.requires_grad_(i == t); head.bias.requires_grad_(i == t)
        self.lora_B_k = nn.ParameterList([nn.Parameter(self.B_cur[i, 0]) for i in range(L)])
        self.lora_B_v = nn.ParameterList([nn.Parameter(self.B_cur[i, 1]) for i in range(L)])
        head = self._network.classifier_pool[t]
        self.params = list(self.lora_B_k) + list(self.lora_B_v) + [head.weight, head.bias]
        if A is not None:
            self.set_A(A)

    def set_A(self, A: torch.Tensor):
        self.A_cur.copy_(A.to(self.engine.dev).reshape(self.A_cur.shape))
        self.engine.lora.set_A(self.A_cur)

    @torch.no_grad()
    def input_matrices(self, loader) -> torch.Tensor:
        """`self._network(inputs, get_cur_feat=True)` over a loader (InfLoRA.py:146-152): per block the token mean of h h^T, h = norm1(x)."""
        eng = self.engine
        eng.lora_merge()
        eng.input_matrix_begin()
        for batch in loader:
            eng.forward(self._images(batch["image"] if isinstance(batch, dict) else batch), None, save=False)
        return eng.input_matrix_end()

    @torch.no_grad()
    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        """InfLoRA.py:104-185: lora_A of the new task = top-r left singular basis of the task's input matrix, outside ('remove') or inside ('retain')
        the subspace kept from the earlier tasks."""
        self.start_task()
        cur = self.input_matrices(train_loader)
        L, r = self.engine.depth, self.rank
        A = torch.empty(L, 2, r, DIM, device=self.engine.dev)
        for i in range(L):
            m = cur[i]
            if self._cur_task > 0:
                inside = self.feature_mat[i].to(m) @ m
                m = m - inside if self.project_type[i] == "remove" else inside
            U, _, _ = torch.linalg.svd(m, full_matrices=False)
            A[i, 0] = A[i, 1] = U[:, :r].T / math.sqrt(3)
        self.set_A(A)

    @torch.no_grad()
    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        """InfLoRA.py:187-213."""
        mats = self.input_matrices(train_loader).cpu().numpy()
        dualgpm_update_v1(list(mats), self.feature_list, self.project_type, self._cur_task, self.total_sessions, self.lame, self.lamb)
        self.feature_mat = [torch.from_numpy(np.dot(f, f.transpose()).astype(np.float32)) for f in self.feature_list]

    # ---- the step ------------------------------------------------------------------------------------
    def _launch_step(self, x, y, clip: bool = True):
        """y: labels relative to the task (observe subtracts `_known_classes`, InfLoRA.py:72)."""
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        bb = self._batch_bufs(B)
        c, t = self._network.class_num, self._cur_task
        lo = t * c
        eng.lora_merge()
        ws = eng.forward(x, None, save=True)
        feat = eng.pooled(ws, 0)
        C = self.total_heads
        check(lib.lc_linear_head(feat.data_ptr(), self.heads_W[lo].data_ptr(), self.heads_b[t].data_ptr(), B, c, DIM, bb["logits"][:, lo:].data_ptr(), C, st), "linear_head")
        torch.add(y, lo, out=bb["yrel"])                          # column index of the target inside the [all heads] logits row
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, bb["yrel"].data_ptr(), B, lo, lo + c, None, 0.0, bb["dlogits"].data_ptr(), bb["pred"].data_ptr(),
                                    self.scal.data_ptr(), st), "loss_ce_masked")
        check(lib.lc_linear_head_backward(bb["dlogits"][:, lo:].data_ptr(), C, feat.data_ptr(), self.heads_W[lo].data_ptr(), c, B, DIM,
                                          self.theta_grad[self.oW + lo * DIM:].data_ptr(), self.theta_grad[self.ob + t * self.cp:].data_ptr(), bb["dfeat"].data_ptr(), st),
              "linear_head_backward")
        eng.backward_tokens(ws, bb["dfeat"], 0, to_tokens=False)
        eng.launches += 3
        return bb

    def observe(self, data):
        x = self._images(data["image"])
        y = (data["label"].to(self.engine.dev, torch.int64) - self._known_classes).contiguous()
        bb = self._launch_step(x, y)
        acc = float(self.scal[1].item()) / x.shape[0]
        return bb["pred"] - self._cur_task * self._network.class_num, acc, _FlatLoss.apply(self, self.scal[0], *self.params)

    @torch.no_grad()
    def inference(self, data):
        """`SiNet_vit.interface` (SiNet.py:121-133): the heads of every task seen so far, concatenated."""
        x = self._images(data["image"])
        y = data["label"].to(self.engine.dev, torch.int64).contiguous()
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        bb = self._batch_bufs(B)
        eng.lora_merge()
        ws = eng.forward(x, None, save=False)
        feat = eng.pooled(ws, 0)
        C, n = self.total_heads, self._network.numtask * self._network.class_num
        bias = self.heads_b[:self._network.numtask, :self._network.class_num].reshape(-1).contiguous()     # the padded slots gathered back to back
        check(lib.lc_linear_head(feat.data_ptr(), self.heads_W.data_ptr(), bias.data_ptr(), B, n, DIM, bb["logits"].data_ptr(), C, st), "linear_head")
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, y.data_ptr(), B, 0, n, None, 0.0, bb["dlogits"].data_ptr(), bb["pred"].data_ptr(),
                                    self.scal.data_ptr(), st), "argmax")
        eng.launches += 2
        return bb["pred"], host_acc(self, self.scal[1], B)


def dualgpm_update_v1(mat_list, feature_list: List[np.ndarray], project_type: List[str], cur_task: int, total_sessions: int, lame: float, lamb: float):
    """`InfLoRA.update_DualGPM` (InfLoRA.py:215-307), in place on the two lists.  Per block: after the first task keep the leading left singular
    vectors of the input matrix up to the energy threshold ('remove', or 'retain' if that is already half the space); afterwards grow a 'remove' basis
    with directions of the residual until the captured energy reaches the threshold, or shrink a 'retain' basis while the energy inside stays above
    1 - threshold; finally swap any 'remove' basis wider than half the dimension for its orthogonal complement."""
    threshold = (lame - lamb) * cur_task / total_sessions + lamb
    if len(feature_list) == 0:
        for act in mat_list:
            U, S, _ = np.linalg.svd(act, full_matrices=False)
            ratio = S ** 2 / (S ** 2).sum()
            r = int(np.sum(np.cumsum(ratio) < threshold))
            feature_list.append(U[:, :max(r, 1)])
            project_type.append("remove" if r < act.shape[0] / 2 else "retain")
    else:
        for i, act in enumerate(mat_list):
            F_i = feature_list[i]
            total = (np.linalg.svd(act, compute_uv=False) ** 2).sum()
            inside = np.dot(np.dot(F_i, F_i.transpose()), act)
            if project_type[i] == "remove":
                U, S, _ = np.linalg.svd(act - inside, full_matrices=False)
                ratio = S ** 2 / total
                acc = (total - (S ** 2).sum()) / total
                r = 0
                while r < ratio.shape[0] and acc < threshold:
                    acc += ratio[r]
                    r += 1
                if r == 0:
                    continue
                grown = np.hstack((F_i, U[:, :r]))
                feature_list[i] = grown[:, :grown.shape[0]] if grown.shape[1] > grown.shape[0] else grown
            else:
                U, S, _ = np.linalg.svd(inside, full_matrices=False)
                ratio = S ** 2 / total
                acc = (S ** 2).sum() / total
                r = 0
                while r < ratio.shape[0] and acc >= 1 - threshold:
                    acc -= ratio[r]
                    r += 1
                if r == 0:
                    continue
                shrunk = F_i - np.dot(np.dot(U[:, :r], U[:, :r].transpose()), F_i)
                U2, _, _ = np.linalg.svd(shrunk)
                feature_list[i] = U2[:, :F_i.shape[1] - r]
    for i, F_i in enumerate(feature_list):
        if project_type[i] == "remove" and F_i.shape[1] > F_i.shape[0] / 2:
            U, _, _ = np.linalg.svd(F_i)
            feature_list[i] = U[:, F_i.shape[1]:]
            project_type[i] = "retain"
        elif project_type[i] == "retain":
            assert F_i.shape[1] <= F_i.shape[0] / 2
