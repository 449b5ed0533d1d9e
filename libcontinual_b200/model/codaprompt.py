"""CodaPrompt on ViT-B/16 — mirror of the reference plugin surface (core/model/codaprompt.py:57-121; pool: core/model/backbone/prompt.py:37-220) on top
of `ViTEngine`; the plugin body is shared with DualPrompt (`_PrefixPromptMethod`).

    model = CodaPrompt(backbone, 768, 100, device=dev, task_num=10, init_cls_num=10, inc_cls_num=10, prompt_length=8, pool_size=100, mu=0.0)

Per block 0-4 the prefix keys / values of every image are the attention-weighted sum of the pool components, alpha[b][k] = cos(q_b * A_k, K_k)
(`lc_coda_prompt_forward`), consumed by the fused prefix attention; the backward maps the prefix-row gradients back to the components, keys and
attention vectors (`lc_coda_prompt_backward`).  Reference quirk kept: `process_task_count` is never called, so every task trains the first
pool_size / n_tasks components (prompt.py:166-168 with task_count = 0).
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn

from .._lib import check, stream_ptr
from ..vit_engine import DIM
from .dualprompt import _PrefixPromptMethod

CODA_LAYERS = (0, 1, 2, 3, 4)


class CodaPromptPool(nn.Module):
    """`core.model.backbone.prompt.CodaPrompt` (prompt.py:37-140): parameters + the Gram-Schmidt (re)initialisation of the current task's components."""

    def __init__(self, emb_d, n_tasks, prompt_param, key_dim=DIM):
        super().__init__()
        self.task_count = 0
        self.emb_d, self.key_d, self.n_tasks = emb_d, key_dim, n_tasks
        self.e_pool_size, self.e_p_length, self.ortho_mu = int(prompt_param[0]), int(prompt_param[1]), prompt_param[2]
        self.e_layers = list(CODA_LAYERS)
        for e in self.e_layers:
            p = nn.init.uniform_(torch.empty(self.e_pool_size, self.e_p_length, emb_d))
            k = nn.init.uniform_(torch.empty(self.e_pool_size, key_dim))
            a = nn.init.uniform_(torch.empty(self.e_pool_size, key_dim))
            setattr(self, f"e_p_{e}", nn.Parameter(self.gram_schmidt(p)))
            setattr(self, f"e_k_{e}", nn.Parameter(self.gram_schmidt(k)))
            setattr(self, f"e_a_{e}", nn.Parameter(self.gram_schmidt(a)))

    def components(self):
        """(s, f): the component range of the current task (prompt.py:166-168)."""
        pt = int(self.e_pool_size / self.n_tasks)
        return int(self.task_count * pt), int((self.task_count + 1) * pt)

    @torch.no_grad()
    def gram_schmidt(self, vv: torch.Tensor) -> torch.Tensor:
        """prompt.py:98-144: rows [0, s) are kept, rows [s, f) become an orthonormal set drawn from N(0, 1) (orthogonalised against ALL earlier rows,
        in order, then normalised), rows >= f are zero.  Same draws from the torch RNG as the reference (one `randn` per new row)."""
        shape = vv.shape
        flat = vv.reshape(shape[0], -1)
        out = torch.zeros_like(flat)
        s, f = self.components()
        out[:s] = flat[:s]
        for k in range(s, f):
            while True:
                vk = torch.randn_like(flat[k])
                acc, degenerate = torch.zeros_like(vk), False
                for j in range(k):
                    den = (out[j] * out[j]).sum()
                    if den < 1e-8:
                        degenerate = True
                        break
                    acc = acc + (vk * out[j]).sum() / den * out[j]
                if not degenerate:
                    out[k] = vk - acc
                    break
        for k in range(s, f):
            out[k] = out[k] / out[k].norm()
        return out.reshape(shape)


class CodaPrompt(_PrefixPromptMethod):
    flag = "coda"

    def _make_pool(self, kwargs):
        if float(kwargs.get("mu", 0.0)) > 0:
            raise NotImplementedError("ortho penalty (mu > 0) is not on the CUDA path; the shipped recipe uses mu = 0.0 (config/codaprompt.yaml)")
        return CodaPromptPool(DIM, kwargs["task_num"], [kwargs["pool_size"], kwargs["prompt_length"], kwargs["mu"]])

    def _pool_tensors(self, pool):
        t = []
        for e in CODA_LAYERS:
            t += [getattr(pool, f"e_p_{e}"), getattr(pool, f"e_k_{e}"), getattr(pool, f"e_a_{e}")]
        return t

    def _prefix_shapes(self):
        return {l: self.pool.e_p_length // 2 for l in CODA_LAYERS}

    def _setup_pool_pointers(self):
        pool, n = self.pool, len(CODA_LAYERS)
        Arr = ctypes.c_void_p * n
        get = lambda nm: [getattr(pool, f"{nm}_{e}") for e in CODA_LAYERS]
        self._K, self._A, self._p = Arr(*[t.data_ptr() for t in get("e_k")]), Arr(*[t.data_ptr() for t in get("e_a")]), Arr(*[t.data_ptr() for t in get("e_p")])
        self._dK = Arr(*[self._grad_view(t).data_ptr() for t in get("e_k")])
        self._dA = Arr(*[self._grad_view(t).data_ptr() for t in get("e_a")])
        self._dp = Arr(*[self._grad_view(t).data_ptr() for t in get("e_p")])
        self._Arr = Arr
        self._coda = {}

    def _coda_bufs(self, B):
        if B not in self._coda:
            dev = self.engine.dev
            s, f = self.pool.components()
            z = lambda: torch.zeros(len(CODA_LAYERS), B, f, device=dev)
            self._coda[B] = dict(alpha=z(), vnorm=z(), dalpha=z(), q=torch.zeros(B, DIM, device=dev))
        return self._coda[B]

    def _prefixes(self, x, bb, train: bool):
        """Query pass + attention-weighted prompt of blocks 0-4 (train and inference use the same components: task_count stays 0)."""
        eng, lib, st, pool = self.engine, self.engine.lib, stream_ptr(), self.pool
        B = x.shape[0]
        cb = self._coda_bufs(B)
        ws1 = eng.forward(x, None, save=False)
        cb["q"].copy_(eng.pooled(ws1, 0))               # the pooled-feature buffer is reused by the prefix-tuned pass
        s, f = pool.components()
        pk = self._Arr(*[bb["prefix"][l][0].data_ptr() for l in CODA_LAYERS]); pv = self._Arr(*[bb["prefix"][l][1].data_ptr() for l in CODA_LAYERS])
        check(lib.lc_coda_prompt_forward(cb["q"].data_ptr(), self._K, self._A, self._p, pk, pv, len(CODA_LAYERS), B, f, pool.e_p_length, DIM,
                                         cb["alpha"].data_ptr(), cb["vnorm"].data_ptr(), st), "coda_prompt_forward")
        if train:
            self.prompt_loss.zero_()
        eng.launches += 2
        return bb["prefix"]

    def _prompt_backward(self, ws, bb):
        eng, lib, st, pool = self.engine, self.engine.lib, stream_ptr(), self.pool
        B = ws.B
        cb = self._coda_bufs(B)
        s, f = pool.components()
        half = pool.e_p_length // 2
        dpk = self._Arr(*[ws.prefix_grads(l, half)[0].data_ptr() for l in CODA_LAYERS]); dpv = self._Arr(*[ws.prefix_grads(l, half)[1].data_ptr() for l in CODA_LAYERS])
        check(lib.lc_coda_prompt_backward(cb["q"].data_ptr(), self._K, self._A, self._p, dpk, dpv, self._dK, self._dA, self._dp, len(CODA_LAYERS), B, f,
                                          pool.e_p_length, DIM, cb["alpha"].data_ptr(), cb["vnorm"].data_ptr(), cb["dalpha"].data_ptr(), st), "coda_prompt_backward")
        eng.launches += 2
