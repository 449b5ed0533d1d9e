"""DualPrompt on ViT-B/16 — mirror of the reference plugin surface (core/model/dualprompt.py:46-128; pool: core/model/backbone/prompt.py:231-337;
ViT prompt branch: core/model/backbone/vit.py:121-131, transformer.py:2263-2296) on top of `ViTEngine`.

    backbone = vit_pt_imnet(pretrained=False, state=<VisionTransformer state_dict>)
    model    = DualPrompt(backbone, 768, 100, device=dev, task_num=10, init_cls_num=10, inc_cls_num=10, g_prompt_length=6, e_prompt_length=20)
    pred, acc, loss = model.observe(batch); optimizer.zero_grad(); loss.backward(); optimizer.step()          # trainer.py:601-606

One step = the no-grad query pass (cls feature), the key match of the three e-prompt layers (task-id bootstrap), the prefix-tuned pass (g-prompts
as 3 + 3 prefix keys / values on blocks 0-1, the task's e-prompt as 10 + 10 on blocks 2-4, inside the fused attention kernel), the masked CE and the
backward down to the first attention, where the prefix-row gradients are summed over the batch into the prompt parameters.  Trainables (g / e
prompts, keys, classifier) live in one flat arena `theta` in the reference's `get_parameters` order; `libcontinual_b200.optim.Adam` updates it.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional  # noqa: F401

import torch
import torch.nn as nn

from .._lib import LcError, check, stream_ptr, host_acc
from ..vit_engine import DIM
from .inflora import _FlatLoss
from .l2p import ViTZoo

G_LAYERS, E_LAYERS = (0, 1), (2, 3, 4)


class DualPromptPool(nn.Module):
    """`core.model.backbone.prompt.DualPrompt` (prompt.py:231-263): parameters only (uniform(0, 1) init, `tensor_prompt`); selection and
    gather run in lc_prompt_key_match / lc_gather_rows_bf16."""

    def __init__(self, emb_d, n_tasks, prompt_param, key_dim=DIM):
        super().__init__()
        self.task_count = 0
        self.emb_d, self.key_d, self.n_tasks = emb_d, key_dim, n_tasks
        self.top_k, self.task_id_bootstrap = 1, True
        self.g_layers, self.e_layers = list(G_LAYERS), list(E_LAYERS)
        self.e_pool_size, self.e_p_length, self.g_p_length = int(prompt_param[0]), int(prompt_param[1]), int(prompt_param[2])
        for g in self.g_layers:
            setattr(self, f"g_p_{g}", nn.Parameter(nn.init.uniform_(torch.empty(self.g_p_length, emb_d))))
        for e in self.e_layers:
            setattr(self, f"e_p_{e}", nn.Parameter(nn.init.uniform_(torch.empty(self.e_pool_size, self.e_p_length, emb_d))))
            setattr(self, f"e_k_{e}", nn.Parameter(nn.init.uniform_(torch.empty(self.e_pool_size, key_dim))))

    def process_task_count(self):
        self.task_count += 1


class _Classifier(nn.Module):
    """nn.Linear whose weight / bias are the first `out_features` rows of the owner's head arena (the reference re-allocates a larger nn.Linear every
    task and copies the old rows: dualprompt.py:64-71)."""

    def __init__(self, W: torch.Tensor, b: torch.Tensor, out_features: int):
        super().__init__()
        self.in_features, self.out_features = W.shape[1], out_features
        self.weight = nn.Parameter(W[:out_features])
        self.bias = nn.Parameter(b[:out_features])

    def _apply(self, fn, recurse=True):
        return self


class Model(nn.Module):
    def __init__(self, backbone, feat_dim, num_class):
        super().__init__()
        self.backbone, self.feat_dim, self.num_class = backbone, feat_dim, num_class
        self.classifier = None


class _PrefixPromptMethod(nn.Module):
    """Shared plugin body of the prefix-tuning methods (dualprompt.py:46-128 and codaprompt.py:57-121 are the same class around different pools):
    flat arena of [pool parameters | classifier], growing classifier, masked CE, `observe` / `inference`.  Subclasses provide the pool and the two
    pool-specific launches: `_prefixes` (query -> per-image BF16 prefix rows of blocks 0-4) and `_prompt_backward` (prefix-row gradients -> pool)."""

    flag = ""

    def _make_pool(self, kwargs) -> nn.Module:
        raise NotImplementedError

    def _pool_tensors(self, pool) -> List[nn.Parameter]:
        raise NotImplementedError

    def _prefix_shapes(self) -> Dict[int, int]:
        """block index -> number of prefix keys (= values)."""
        raise NotImplementedError

    def __init__(self, backbone: ViTZoo, feat_dim, num_class, **kwargs):
        super().__init__()
        if not isinstance(backbone, ViTZoo):
            raise LcError(f"{type(self).__name__} needs a libcontinual_b200 ViTZoo backbone")
        assert feat_dim == DIM
        self.kwargs = kwargs
        self.device = torch.device(kwargs.get("device", backbone.engine.dev))
        self.backbone = backbone
        self.engine = eng = backbone.engine
        self.network = Model(backbone, feat_dim, kwargs["init_cls_num"])
        pool = self._make_pool(kwargs)
        backbone.prompt, backbone.prompt_flag = pool, self.flag
        self.pool = pool
        self.num_class = num_class
        self.task_idx = 0
        self.last_out_dim = 0
        self.out_dim = 0
        dev = eng.dev
        # flat arena in the order of `list(prompt.parameters()) + list(classifier.parameters())` (dualprompt.py:127-128)
        tensors = self._pool_tensors(pool)
        self.n_prompt = sum(t.numel() for t in tensors)
        self.oW, self.ob = self.n_prompt, self.n_prompt + num_class * DIM
        self.theta = torch.zeros(self.ob + num_class, device=dev)
        self.theta_grad = torch.zeros_like(self.theta)
        off = 0
        for t in tensors:
            n = t.numel()
            self.theta[off:off + n].copy_(t.detach().reshape(-1))
            t.data = self.theta[off:off + n].view(t.shape)
            off += n
        self.prompt_params = tensors
        self.head_W = self.theta[self.oW:self.ob].view(num_class, DIM)
        self.head_b = self.theta[self.ob:]
        first = nn.Linear(DIM, kwargs["init_cls_num"])                       # `Model.__init__` (dualprompt.py:43): the initial classifier
        self.head_W[:kwargs["init_cls_num"]].copy_(first.weight.detach()); self.head_b[:kwargs["init_cls_num"]].copy_(first.bias.detach())
        self.network.classifier = _Classifier(self.head_W, self.head_b, kwargs["init_cls_num"])
        self.autograd_grads: Optional[torch.Tensor] = None
        self._bufs: Dict[int, dict] = {}
        self.scal = torch.zeros(8, device=dev)
        self.prompt_loss = torch.zeros(1, device=dev)
        self._setup_pool_pointers()

    def _setup_pool_pointers(self):
        pass

    # ---- arena helpers -----------------------------------------------------------------------------
    def _grad_view(self, p: torch.Tensor, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        off = (p.data_ptr() - self.theta.data_ptr()) // 4
        return (self.theta_grad if arena is None else arena)[off:off + p.numel()].view(p.shape)

    def trainable_params(self) -> List[nn.Parameter]:
        return self.prompt_params + [self.network.classifier.weight, self.network.classifier.bias]

    def get_parameters(self, config):
        return self.trainable_params()

    def _batch_bufs(self, B):
        if B not in self._bufs:
            dev, C = self.engine.dev, self.num_class
            bf = torch.bfloat16
            pre = {l: (torch.zeros(B, n, DIM, device=dev, dtype=bf), torch.zeros(B, n, DIM, device=dev, dtype=bf)) for l, n in self._prefix_shapes().items()}
            self._bufs[B] = dict(logits=torch.zeros(B, C, device=dev), dlogits=torch.zeros(B, C, device=dev), pred=torch.zeros(B, dtype=torch.int64, device=dev),
                                 dfeat=torch.zeros(B, DIM, device=dev), idx=torch.zeros(len(E_LAYERS), B, dtype=torch.int64, device=dev), prefix=pre)
        return self._bufs[B]

    def _to_device(self, data):
        x = data["image"].to(self.engine.dev, torch.float32, non_blocking=True).contiguous()
        y = data["label"].to(self.engine.dev, torch.int64, non_blocking=True).contiguous()
        return x, y

    # ---- plugin surface ------------------------------------------------------------------------------
    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.task_idx = task_idx
        self.backbone.task_id = task_idx
        old = self.network.classifier.out_features
        new = self.kwargs["init_cls_num"] + task_idx * self.kwargs["inc_cls_num"]
        if new > old:
            fresh = nn.Linear(DIM, new)                                      # same RNG draws as the reference's `new_fc` (dualprompt.py:66)
            with torch.no_grad():
                self.head_W[old:new].copy_(fresh.weight[old:new]); self.head_b[old:new].copy_(fresh.bias[old:new])
        self.network.classifier = _Classifier(self.head_W, self.head_b, new)
        self.out_dim = new

    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        self.last_out_dim = self.out_dim

    def _launch_step(self, x, y, clip: bool = True):
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        bb = self._batch_bufs(B)
        self.theta_grad.zero_()                  # rows of other tasks' prompts / old classes keep exact zeros (Adam then leaves them untouched)
        prefix = self._prefixes(x, bb, train=True)
        ws = eng.forward(x, None, save=True, prefix=prefix)
        feat = eng.pooled(ws, 0)
        C, n = self.num_class, self.out_dim
        eng.linear_head(feat, self.head_W[:n], self.head_b[:n], bb["logits"])
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, y.data_ptr(), B, self.last_out_dim, n, self.prompt_loss.data_ptr(), 1.0, bb["dlogits"].data_ptr(),
                                    bb["pred"].data_ptr(), self.scal.data_ptr(), st), "loss_ce_masked")
        lo = self.last_out_dim
        gW = self.theta_grad[self.oW + lo * DIM:]
        gb = self.theta_grad[self.ob + lo:]
        check(lib.lc_linear_head_backward(bb["dlogits"][:, lo:].data_ptr(), C, feat.data_ptr(), self.head_W[lo].data_ptr(), n - lo, B, DIM, gW.data_ptr(),
                                          gb.data_ptr(), bb["dfeat"].data_ptr(), st), "linear_head_backward")
        eng.backward_tokens(ws, bb["dfeat"], 0, to_tokens=False)
        self._prompt_backward(ws, bb)
        eng.launches += 3
        return bb

    def observe(self, data):
        x, y = self._to_device(data)
        bb = self._launch_step(x, y)
        B = x.shape[0]
        acc = float(self.scal[1].item()) / B
        return bb["pred"], acc, _FlatLoss.apply(self, self.scal[0], *self.trainable_params())

    @torch.no_grad()
    def inference(self, data):
        """dualprompt.py:110-122: per-sample top-1 e-prompt (prompt.py:290-292), argmax over every class seen so far."""
        x, y = self._to_device(data)
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        bb = self._batch_bufs(B)
        prefix = self._prefixes(x, bb, train=False)
        ws = eng.forward(x, None, save=False, prefix=prefix)
        feat = eng.pooled(ws, 0)
        C, n = self.num_class, self.network.classifier.out_features
        eng.linear_head(feat, self.head_W[:n], self.head_b[:n], bb["logits"])
        check(lib.lc_loss_ce_masked(bb["logits"].data_ptr(), C, y.data_ptr(), B, 0, n, None, 0.0, bb["dlogits"].data_ptr(), bb["pred"].data_ptr(),
                                    self.scal.data_ptr(), st), "argmax")
        eng.launches += 1
        return bb["pred"], host_acc(self, self.scal[1], B)


class DualPrompt(_PrefixPromptMethod):
    flag = "dual"

    def _make_pool(self, kwargs):
        return DualPromptPool(DIM, kwargs["task_num"], [10, kwargs["e_prompt_length"], kwargs["g_prompt_length"]])

    def _pool_tensors(self, pool):
        tensors = [getattr(pool, f"g_p_{g}") for g in G_LAYERS]
        for e in E_LAYERS:
            tensors += [getattr(pool, f"e_p_{e}"), getattr(pool, f"e_k_{e}")]
        return tensors

    def _prefix_shapes(self):
        d = {l: self.pool.g_p_length // 2 for l in G_LAYERS}
        d.update({l: self.pool.e_p_length // 2 for l in E_LAYERS})
        return d

    def _setup_pool_pointers(self):
        pool = self.pool
        PtrArr = ctypes.c_void_p * len(E_LAYERS)
        self._keys = PtrArr(*[getattr(pool, f"e_k_{e}").data_ptr() for e in E_LAYERS])
        self._dkeys = PtrArr(*[self._grad_view(getattr(pool, f"e_k_{e}")).data_ptr() for e in E_LAYERS])

    def _prefixes(self, x, bb, train: bool):
        """Query pass + key match + gather of the per-image BF16 prefix rows of blocks 0-4."""
        eng, lib, st = self.engine, self.engine.lib, stream_ptr()
        B = x.shape[0]
        pool = self.pool
        ws1 = eng.forward(x, None, save=False)
        q = eng.pooled(ws1, 0)
        check(lib.lc_prompt_key_match(q.data_ptr(), self._keys, self._dkeys if train else None, len(E_LAYERS), B, pool.e_pool_size, DIM,
                                      self.task_idx if train else -1, bb["idx"].data_ptr(), self.prompt_loss.data_ptr() if train else None, st), "prompt_key_match")
        gl, el = pool.g_p_length // 2, pool.e_p_length // 2
        for l in G_LAYERS:
            g = getattr(pool, f"g_p_{l}")
            for half, out in enumerate(bb["prefix"][l]):
                check(lib.lc_gather_rows_bf16(g.data_ptr() + 4 * half * gl * DIM, None, 0, gl, DIM, B, out.data_ptr(), st), "gather g")
        for li, l in enumerate(E_LAYERS):
            e = getattr(pool, f"e_p_{l}")
            for half, out in enumerate(bb["prefix"][l]):
                check(lib.lc_gather_rows_bf16(e.data_ptr() + 4 * half * el * DIM, bb["idx"][li].data_ptr(), pool.e_p_length * DIM, el, DIM, B, out.data_ptr(), st),
                      "gather e")
        eng.launches += 1 + 2 * (len(G_LAYERS) + len(E_LAYERS))
        return bb["prefix"]

    def _prompt_backward(self, ws, bb):
        """Prefix-row gradients summed over the batch into the prompt parameters (`expand(len(x_querry), -1, -1)`, prompt.py:284,306)."""
        eng, lib, st, pool = self.engine, self.engine.lib, stream_ptr(), self.pool
        B = ws.B
        gl, el = pool.g_p_length // 2, pool.e_p_length // 2
        for l in G_LAYERS:
            gg = self._grad_view(getattr(pool, f"g_p_{l}"))
            for half, d in enumerate(ws.prefix_grads(l, gl)):
                check(lib.lc_sum_batch_rows(d.data_ptr(), gl * DIM, B, gl, DIM, gg[half * gl:].data_ptr(), st), "sum g")
        for l in E_LAYERS:
            ge = self._grad_view(getattr(pool, f"e_p_{l}"))[self.task_idx]
            for half, d in enumerate(ws.prefix_grads(l, el)):
                check(lib.lc_sum_batch_rows(d.data_ptr(), el * DIM, B, el, DIM, ge[half * el:].data_ptr(), st), "sum e")
        eng.launches += 2 * (len(G_LAYERS) + len(E_LAYERS))
