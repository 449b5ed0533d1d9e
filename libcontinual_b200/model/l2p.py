"""L2P on ViT-B/16 — mirror of the reference plugin surface (core/model/l2p.py:43-122, core/model/backbone/vit.py:47-138,298-299,
core/model/backbone/prompt.py:346-406) on top of `ViTEngine`.

    backbone = vit_pt_imnet(pretrained=False, state=<VisionTransformer state_dict>)
    model    = L2P(backbone, device, init_cls_num=10, inc_cls_num=10, num_class=100, task_num=10, feat_dim=768, prompt_length=5,
                   pool_size=10, top_k=5, pull_constraint_coeff=1.0)
    optimizer.zero_grad(); pred, acc, loss = model.observe(batch); optimizer.step()          # trainer.py:592-606

`observe` runs the query pass, the pool selection, the prompted pass, the masked loss and the backward down to the prompt rows, clips the
gradient norm to 1.0 (l2p.py:104) and leaves the result in `.grad` of the four trainable tensors, exactly what the reference's
`loss.backward(); clip_grad_norm_` leaves behind.  The trainables live in one flat arena (`theta`), so `libcontinual_b200.optim.Adam` updates
them with one kernel; `torch.optim.Adam` on `get_parameters()` works as well.
"""
from __future__ import annotations

import math
import os
from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import _lib
from .._lib import check, stream_ptr, host_acc
from ..vit_engine import DIM, ViTEngine, vit_param_layout

CLIP_SCRATCH_FLOATS = 2 * 296 + 8


class PromptPool(nn.Module):
    """`core.model.backbone.prompt.L2P` (prompt.py:346-367): parameters only; selection runs in lc_l2p_select."""

    def __init__(self, length, pool_size, top_k, num_layers=1, embed_dim=DIM, prompt_init=nn.init.uniform_, **_):
        super().__init__()
        assert num_layers == 1, "L2P prepends prompts at layer 0 only (transformer.py:2233-2235)"
        self.length, self.pool_size, self.top_k, self.num_layers, self.embed_dim = length, pool_size, top_k, num_layers, embed_dim
        self.prompt_init = prompt_init
        self.prompt = nn.Parameter(torch.empty(num_layers, pool_size, length, embed_dim))
        self.prompt_key = nn.Parameter(torch.empty(pool_size, embed_dim))
        prompt_init(self.prompt)
        prompt_init(self.prompt_key)


class ViTZoo(nn.Module):
    """Frozen ViT-B/16 (`ViTZoo`, vit.py:47-138).  The backbone weights live in the engine (fp32 master + BF16 GEMM copies), not as
    nn.Parameters: they never receive gradients (l2p.py:66-71)."""

    def __init__(self, pretrained: bool = False, model_name: str = "vit_base_patch16_224", state: Optional[Dict[str, torch.Tensor]] = None,
                 device="cuda:0", depth: int = 12, **kwargs):
        super().__init__()
        self.task_id = None
        self.feat_dim = DIM
        self.depth = depth
        self.attn_layer = kwargs.get("attn_layer", "MultiHeadAttention")       # vit.py:48; the adapters themselves are installed by the method plugin
        self.lora_rank = int(kwargs.get("lora_rank", 10))
        self.engine = ViTEngine(depth=depth, device=device)
        if state is None:
            if pretrained:
                # the reference asks timm to download `model_name` (vit.py:67); here the checkpoint must already be on disk
                ckpt = os.environ.get("LC_B200_VIT_CHECKPOINT")
                if not ckpt or not os.path.exists(ckpt):
                    raise _lib.LcError(f"pretrained=True ({model_name}): point LC_B200_VIT_CHECKPOINT at a local VisionTransformer / timm state_dict "
                                       "(torch.save file), or pass state=; this build has no network access and does not depend on timm")
                state = torch.load(ckpt, map_location="cpu")
                state = state.get("state_dict", state.get("model", state)) if isinstance(state, dict) else state
            else:
                state = self._random_state(depth)
        self.load_backbone_state(state)
        self.prompt = None
        self.prompt_flag = ""

    @staticmethod
    def _random_state(depth):
        """`_init_weights` (transformer.py:2207-2214): trunc_normal(0.02) matrices, zero biases, unit LayerNorm."""
        st = {}
        for name, shape in vit_param_layout(depth):
            if name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name == "norm.weight":
                st[name] = torch.ones(shape)
            elif name.endswith(".bias"):
                st[name] = torch.zeros(shape)
            elif name == "patch_embed.proj.weight":
                t = torch.empty(shape)
                nn.init.kaiming_uniform_(t, a=math.sqrt(5))
                st[name] = t
            else:
                st[name] = nn.init.trunc_normal_(torch.empty(shape), std=0.02)
        return st

    def load_backbone_state(self, state: Dict[str, torch.Tensor]):
        """Accepts the reference `VisionTransformer` keys or timm's (`.norm1.`/`.norm2.`/`blocks.` are mapped as in vit.py:70-84)."""
        mapped = {}
        for k, v in state.items():
            nk = k
            if ".norm1." in nk:
                nk = nk.replace(".norm1.", ".ln_1.")
            if ".norm2." in nk:
                nk = nk.replace(".norm2.", ".ln_2.")
            if nk.startswith("blocks."):
                nk = "transformer." + nk
            mapped[nk] = v
        self.engine.load_state(mapped)

    def create_prompt(self, prompt_flag, **kwargs):
        self.prompt_flag = prompt_flag
        if prompt_flag != "l2p":
            raise NotImplementedError(f"prompt_flag={prompt_flag!r}: only 'l2p' is on the CUDA path so far")
        self.prompt = PromptPool(**kwargs)


def vit_pt_imnet(pretrained=False, **kwargs):
    """vit.py:298-299."""
    return ViTZoo(pretrained, **kwargs)


class Model(nn.Module):
    """l2p.py:31-40: backbone + nn.Linear(embed_dim, total_cls_num)."""

    def __init__(self, backbone, embed_dim, total_cls_num):
        super().__init__()
        self.backbone = backbone
        self.classifier = nn.Linear(embed_dim, total_cls_num, bias=True)


class L2P(nn.Module):
    def __init__(self, backbone: ViTZoo, device, **kwargs):
        super().__init__()
        self.device = torch.device(device)
        self.init_cls_num = kwargs["init_cls_num"]
        self.inc_cls_num = kwargs["inc_cls_num"]
        self.total_cls_num = kwargs["num_class"]
        self.task_num = kwargs["task_num"]
        self.embed_dim = kwargs["feat_dim"]
        self.pull_constraint_coeff = kwargs["pull_constraint_coeff"]
        # extension (not a reference kwarg): under torch.distributed, vote on the prompt histogram of the GLOBAL batch instead of each rank's shard — exactly
        # the single-GPU result at the global batch size (the reference's own DDP path is dead code; per-rank voting is what plain DDP would do)
        self.sync_vote = bool(kwargs.get("sync_vote", False))
        self.vote_group = kwargs.get("vote_group", None)
        self.cur_task_id = 0
        self._known_classes = 0
        assert self.embed_dim == DIM
        self.engine = backbone.engine
        self.network = Model(backbone, self.embed_dim, self.total_cls_num)
        backbone.create_prompt(prompt_flag="l2p", length=kwargs["prompt_length"], prompt_init=nn.init.uniform_, pool_size=kwargs["pool_size"],
                               top_k=kwargs["top_k"], num_layers=1, embed_dim=self.embed_dim)
        pool = backbone.prompt
        self.pool_size, self.top_k, self.length = pool.pool_size, pool.top_k, pool.length
        self.n_prompt = self.top_k * self.length
        # flat arena of the trainables, in the reference's named_parameters() order: prompt, prompt_key, classifier.weight, classifier.bias
        tensors = [pool.prompt, pool.prompt_key, self.network.classifier.weight, self.network.classifier.bias]
        sizes = [t.numel() for t in tensors]
        assert all(s % 4 == 0 for s in sizes)
        dev = self.engine.dev
        self.theta = torch.empty(sum(sizes), device=dev)
        self.theta_grad = torch.zeros_like(self.theta)
        self.offsets = []
        off = 0
        for t, n in zip(tensors, sizes):
            self.theta[off:off + n].copy_(t.detach().reshape(-1))
            t.data = self.theta[off:off + n].view(t.shape)
            self.offsets.append(off)
            off += n
        self.unfrezeed_params = tensors                      # (sic) reference attribute name, l2p.py:66
        for t in tensors:
            t.requires_grad_(True)
        self._grad_views = [self.theta_grad[o:o + n].view(t.shape) for t, o, n in zip(tensors, self.offsets, sizes)]
        # small device buffers
        P = self.pool_size
        self._bufs: Dict[int, dict] = {}
        self.ids = torch.zeros(self.top_k, dtype=torch.int64, device=dev)
        self.hist = torch.zeros(P, dtype=torch.int32, device=dev)
        self.reduce_sim = torch.zeros(1, device=dev)
        self.dkey_raw = torch.zeros(P, DIM, device=dev)
        self.sel_scratch = torch.zeros(DIM, device=dev)
        self.prompts = torch.zeros(self.n_prompt, DIM, device=dev)
        self.dprompts = torch.zeros(self.n_prompt, DIM, device=dev)
        self.scal = torch.zeros(8, device=dev)
        self.clip_scratch = torch.zeros(CLIP_SCRATCH_FLOATS, device=dev)
        self.grad_norm = torch.zeros(1, device=dev)

    # ---- views ---------------------------------------------------------------------------------
    def _view(self, i, arena=None):
        t = self.unfrezeed_params[i]
        a = self.theta if arena is None else arena
        return a[self.offsets[i]:self.offsets[i] + t.numel()]

    def _batch_bufs(self, B):
        if B not in self._bufs:
            dev = self.engine.dev
            self._bufs[B] = dict(sim=torch.zeros(B, self.pool_size, device=dev), logits=torch.zeros(B, self.total_cls_num, device=dev),
                                 dlogits=torch.zeros(B, self.total_cls_num, device=dev), pred=torch.zeros(B, dtype=torch.int64, device=dev),
                                 dfeat=torch.zeros(B, DIM, device=dev))
        return self._bufs[B]

    def _vote_is_global(self) -> bool:
        import torch.distributed as dist
        return self.sync_vote and dist.is_available() and dist.is_initialized() and dist.get_world_size(self.vote_group) > 1

    # ---- plugin surface --------------------------------------------------------------------------
    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.cur_task_id = task_idx

    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        self._known_classes += self.init_cls_num if task_idx == 0 else self.inc_cls_num

    def get_parameters(self, config):
        return self.unfrezeed_params

    def _to_device(self, data):
        x = data["image"].to(self.engine.dev, torch.float32, non_blocking=True).contiguous()
        y = data["label"].to(self.engine.dev, torch.int64, non_blocking=True).contiguous()
        return x, y

    def _forward_logits(self, x, save):
        """`Model.forward` (l2p.py:38-40) = `ViTZoo.forward` l2p branch (vit.py:102-119) + classifier."""
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        bb = self._batch_bufs(B)
        ws1 = eng.forward(x, None, save=False)
        q = eng.pooled(ws1, 0)
        sel = (q.data_ptr(), self._view(1).data_ptr(), B, self.pool_size, DIM, self.top_k, bb["sim"].data_ptr(), self.ids.data_ptr(), self.hist.data_ptr(),
               self.reduce_sim.data_ptr(), self.dkey_raw.data_ptr(), self.sel_scratch.data_ptr())
        if self._vote_is_global():
            # data parallel with `sync_vote`: the majority vote (prompt.py:380-401) is taken over the GLOBAL batch — per-rank counts, one SUM all-reduce of
            # the [pool] int32 histogram (capturable, like the gradient all-reduce), then the vote / pull constraint on the summed histogram
            from ..parallel import allreduce_sum_
            check(lib.lc_l2p_select_phase(*sel, 1, st), "l2p_select(counts)")
            allreduce_sum_(self.hist, self.vote_group)
            check(lib.lc_l2p_select_phase(*sel, 2, st), "l2p_select(vote)")
            eng.launches += 2      # the all-reduce + one more launch of the vote kernel
        else:
            check(lib.lc_l2p_select(*sel, st), "l2p_select")
        check(lib.lc_l2p_gather(self._view(0).data_ptr(), self.ids.data_ptr(), self.prompts.data_ptr(), 1, self.top_k, self.length, DIM, st), "l2p_gather")
        ws2 = eng.forward(x, self.prompts, save=save)
        feat = eng.pooled(ws2, self.n_prompt)
        eng.linear_head(feat, self._view(2).view(self.total_cls_num, DIM), self._view(3), bb["logits"])
        eng.launches += 3          # l2p_sim + l2p_select, l2p_gather (the head counts its own)
        return ws2, feat, bb

    def _launch_step(self, x, y, clip: bool = True):
        """Everything `observe` puts on the stream (no host synchronisation): capturable into a CUDA graph.  clip=False stops before
        `clip_grad_norm_` so that a data-parallel caller can average `theta_grad` across ranks first (DDP averages during backward,
        the clip at l2p.py:104 then sees the averaged gradient)."""
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        lo = 0 if self.cur_task_id == 0 else self._known_classes
        hi = lo + (self.init_cls_num if self.cur_task_id == 0 else self.inc_cls_num)
        ws2, feat, bb = self._forward_logits(x, save=True)
        C = self.total_cls_num
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, y.data_ptr(), B, lo, hi, self.reduce_sim.data_ptr(), -float(self.pull_constraint_coeff),
                                    bb["dlogits"].data_ptr(), bb["pred"].data_ptr(), self.scal.data_ptr(), st), "loss_ce_masked")
        check(lib.lc_linear_head_backward(bb["dlogits"].data_ptr(), C, feat.data_ptr(), self._view(2).data_ptr(), C, B, DIM,
                                          self._view(2, self.theta_grad).data_ptr(), self._view(3, self.theta_grad).data_ptr(), bb["dfeat"].data_ptr(), st),
              "linear_head_backward")
        g = eng.backward_tokens(ws2, bb["dfeat"], self.n_prompt)
        eng.prompt_row_grads(g, self.n_prompt, self.dprompts)
        check(lib.lc_l2p_backward(self.dprompts.data_ptr(), self.ids.data_ptr(), self.pool_size, self.top_k, self.length, DIM,
                                  self._view(0, self.theta_grad).data_ptr(), self.dkey_raw.data_ptr(), -float(self.pull_constraint_coeff),
                                  self._view(1, self.theta_grad).data_ptr(), st), "l2p_backward")
        eng.launches += 3      # loss, head backward, l2p backward
        if clip:
            self._launch_clip()
        return bb

    def _launch_clip(self):
        check(self.engine.lib.lc_clip_grad_norm(self.theta_grad.data_ptr(), self.theta_grad.numel(), 1.0, self.clip_scratch.data_ptr(),
                                                self.grad_norm.data_ptr(), stream_ptr()), "clip_grad_norm")
        self.engine.launches += 2

    def observe(self, data):
        x, y = self._to_device(data)
        bb = self._launch_step(x, y)
        for t, gv in zip(self.unfrezeed_params, self._grad_views):
            t.grad = gv
        B = x.shape[0]
        acc = float(self.scal[1].item()) / B
        return bb["pred"], acc, self.scal[0]

    @torch.no_grad()
    def inference(self, data):
        """l2p.py:109-116: argmax over ALL classes (no task mask at test time)."""
        x, y = self._to_device(data)
        eng, lib = self.engine, self.engine.lib
        B = x.shape[0]
        _, _, bb = self._forward_logits(x, save=False)
        C = self.total_cls_num
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, y.data_ptr(), B, 0, C, None, 0.0, bb["dlogits"].data_ptr(), bb["pred"].data_ptr(),
                                    self.scal.data_ptr(), stream_ptr()), "argmax")
        eng.launches += 1
        return bb["pred"], host_acc(self, self.scal[1], B)
