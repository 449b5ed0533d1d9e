"""ORACLE — TEST INFRASTRUCTURE ONLY.  Never imported by the product path (`libcontinual_b200/`).

CPU restatement (PyTorch fp32, functional style, autograd for derivatives) of the reference's per-step
training hot path (RL-VIG/LibContinual @ a399ed4).  Allowed importers: `tests/`,
`__graft_entry__.smoke()`, `bench.py` (`cpu_baseline` leg and `--impl reference`).

Parity pinning: the reference ships no tests / golden vectors, so this restatement is
pinned against outputs of the reference ITSELF, imported in the build container through
`oracle/ref_shim.py` by `oracle/make_golden.py`; the vectors live in `tests/golden/*.npz` and are
re-checked by `tests/test_oracle_golden.py` on every CPU run.

All `file:line` citations are relative to the reference root.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Tensor = torch.Tensor
BN_EPS = 1e-5       # nn.BatchNorm2d default, used by core/model/backbone/resnet.py:296,299,332
BN_MOMENTUM = 0.1   # nn.BatchNorm2d default


# ----------------------------------------------------------------------------------------------
# cifar_resnet32  (core/model/backbone/resnet.py:289-412, factory :760-763)
# ----------------------------------------------------------------------------------------------
def cifar_resnet_layout(depth: int = 32, in_ch: int = 3) -> Tuple[List[Tuple[str, Tuple[int, ...]]], List[Tuple[str, Tuple[int, ...]]]]:
    """Parameter and buffer (name, shape) lists in the reference's `named_parameters()` /
    `named_buffers()` order (module registration order of resnet.py:334-340, 294-301, 361-376)."""
    assert (depth - 2) % 6 == 0
    nblk = (depth - 2) // 6
    params: List[Tuple[str, Tuple[int, ...]]] = []
    bufs: List[Tuple[str, Tuple[int, ...]]] = []

    def bn(prefix, c):
        params.append((prefix + ".weight", (c,)))
        params.append((prefix + ".bias", (c,)))
        bufs.append((prefix + ".running_mean", (c,)))
        bufs.append((prefix + ".running_var", (c,)))
        bufs.append((prefix + ".num_batches_tracked", ()))

    params.append(("conv_1_3x3.weight", (16, in_ch, 3, 3)))
    bn("bn_1", 16)
    inpl = 16
    for s, planes in enumerate((16, 32, 64), start=1):
        for b in range(nblk):
            pre = f"stage_{s}.{b}"
            cin = inpl if b == 0 else planes
            params.append((pre + ".conv_a.weight", (planes, cin, 3, 3)))
            bn(pre + ".bn_a", planes)
            params.append((pre + ".conv_b.weight", (planes, planes, 3, 3)))
            bn(pre + ".bn_b", planes)
            if b == 0 and (s > 1):
                params.append((pre + ".downsample.0.weight", (planes, cin, 1, 1)))
                bn(pre + ".downsample.1", planes)
        inpl = planes
    return params, bufs


def cifar_resnet_init(rng: np.random.Generator, depth: int = 32, in_ch: int = 3) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    """Same init DISTRIBUTIONS as resnet.py:345-352 (conv ~ N(0, sqrt(2/(k*k*Cout))), BN weight 1 / bias 0),
    drawn from a numpy Generator so that the values are platform-stable for fixtures."""
    pl, bl = cifar_resnet_layout(depth, in_ch)
    params, bufs = {}, {}
    for name, shape in pl:
        if len(shape) == 4:
            n = shape[2] * shape[3] * shape[0]
            params[name] = torch.from_numpy((rng.standard_normal(shape) * math.sqrt(2.0 / n)).astype(np.float32))
        elif name.endswith(".weight"):
            params[name] = torch.ones(shape)
        else:
            params[name] = torch.zeros(shape)
    for name, shape in bl:
        if name.endswith("running_var"):
            bufs[name] = torch.ones(shape)
        elif name.endswith("num_batches_tracked"):
            bufs[name] = torch.zeros((), dtype=torch.int64)
        else:
            bufs[name] = torch.zeros(shape)
    return params, bufs


def _bn(x: Tensor, p: Dict[str, Tensor], b: Dict[str, Tensor], prefix: str, train: bool) -> Tensor:
    # nn.BatchNorm2d forward: batch statistics (biased var) in train mode and in-place running-stat
    # update with momentum 0.1 / unbiased var; running statistics in eval mode.
    if train:
        b[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, b[prefix + ".running_mean"], b[prefix + ".running_var"], p[prefix + ".weight"],
                        p[prefix + ".bias"], training=train, momentum=BN_MOMENTUM, eps=BN_EPS)


def _tf32_rna(t: Tensor) -> Tensor:
    """Round fp32 to TF32 (10 mantissa bits), round-to-nearest with ties away from zero — PTX `cvt.rna.tf32.f32`."""
    i = t.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


class _TF32Conv3x3(torch.autograd.Function):
    """Arithmetic class of the product's tensor-core mode (`precision='tc'`) for the stride-1 3x3 convolutions: forward and data
    gradient multiply TF32-rounded operands (cvt.rna) with fp32 accumulation — what cuDNN does for the reference's nn.Conv2d under
    PyTorch's default `torch.backends.cudnn.allow_tf32 = True`; the weight gradient multiplies BF16-rounded operands
    (round-to-nearest-even) with fp32 accumulation."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return F.conv2d(_tf32_rna(x), _tf32_rna(w), None, 1, 1)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dx = torch.nn.grad.conv2d_input(x.shape, _tf32_rna(w), _tf32_rna(dy), stride=1, padding=1)
        dw = torch.nn.grad.conv2d_weight(x.bfloat16().float(), w.shape, dy.bfloat16().float(), stride=1, padding=1)
        return dx, dw


def _conv3x3(x: Tensor, w: Tensor, stride: int, conv_mode: str) -> Tensor:
    if conv_mode == "tc" and stride == 1 and w.shape[0] == w.shape[1]:
        return _TF32Conv3x3.apply(x, w)
    return F.conv2d(x, w, None, stride, 1)


def cifar_resnet_forward(p: Dict[str, Tensor], b: Dict[str, Tensor], x: Tensor, train: bool, depth: int = 32, conv_mode: str = "fp32",
                         last_relu: bool = True) -> Dict[str, object]:
    """resnet.py:381-395 (network) and :303-316 (basic block).  Returns {'fmaps': [x1, x2, x3], 'features': [B, 64]}.
    conv_mode 'tc' restates the tensor-core arithmetic class for the square stride-1 3x3 layers (see _TF32Conv3x3)."""
    nblk = (depth - 2) // 6
    h = F.conv2d(x, p["conv_1_3x3.weight"], None, 1, 1)
    h = F.relu(_bn(h, p, b, "bn_1", train))
    fmaps = []
    for s in (1, 2, 3):
        for k in range(nblk):
            pre = f"stage_{s}.{k}"
            stride = 2 if (k == 0 and s > 1) else 1
            r = h
            y = _conv3x3(h, p[pre + ".conv_a.weight"], stride, conv_mode)
            y = F.relu(_bn(y, p, b, pre + ".bn_a", train))
            y = _conv3x3(y, p[pre + ".conv_b.weight"], 1, conv_mode)
            y = _bn(y, p, b, pre + ".bn_b", train)
            if (pre + ".downsample.0.weight") in p:
                r = F.conv2d(h, p[pre + ".downsample.0.weight"], None, stride, 0)
                r = _bn(r, p, b, pre + ".downsample.1", train)
            # LUCIR's modified_BasicBlock(last=True) drops the final ReLU of the network (resnet.py:497-502)
            h = (r + y) if (not last_relu and s == 3 and k == nblk - 1) else F.relu(r + y)
        fmaps.append(h)
    feats = F.avg_pool2d(h, 8).flatten(1)          # nn.AvgPool2d(8), resnet.py:340,389-390
    return {"fmaps": fmaps, "features": feats}


# ----------------------------------------------------------------------------------------------
# resnet18  (core/model/backbone/resnet.py:26-64 BasicBlock, :110-246 ResNet, factory :259-267)
# ----------------------------------------------------------------------------------------------
def resnet18_layout(in_ch: int = 3) -> Tuple[List[Tuple[str, Tuple[int, ...]]], List[Tuple[str, Tuple[int, ...]]]]:
    """Parameter / buffer (name, shape) lists in the reference's registration order: stem `conv1` = Sequential(Conv2d, BatchNorm2d, ReLU[, MaxPool2d])
    (resnet.py:133-150), `layerN.M.{conv1,bn1,conv2,bn2[,downsample.0,downsample.1]}` (:196-218, :40-46).  The unused `fc` (:191) is left out."""
    params: List[Tuple[str, Tuple[int, ...]]] = []
    bufs: List[Tuple[str, Tuple[int, ...]]] = []

    def bn(prefix, c):
        params.append((prefix + ".weight", (c,)))
        params.append((prefix + ".bias", (c,)))
        bufs.append((prefix + ".running_mean", (c,)))
        bufs.append((prefix + ".running_var", (c,)))
        bufs.append((prefix + ".num_batches_tracked", ()))

    params.append(("conv1.0.weight", (64, in_ch, 3, 3)))
    bn("conv1.1", 64)
    inpl = 64
    for li, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for b in range(2):
            pre = f"layer{li}.{b}"
            s, cin = (stride, inpl) if b == 0 else (1, planes)
            params.append((pre + ".conv1.weight", (planes, cin, 3, 3)))
            bn(pre + ".bn1", planes)
            params.append((pre + ".conv2.weight", (planes, planes, 3, 3)))
            bn(pre + ".bn2", planes)
            if b == 0 and (s != 1 or cin != planes):
                params.append((pre + ".downsample.0.weight", (planes, cin, 1, 1)))
                bn(pre + ".downsample.1", planes)
        inpl = planes
    return params, bufs


def resnet18_init(rng: np.random.Generator, in_ch: int = 3) -> Tuple[Dict[str, Tensor], Dict[str, Tensor]]:
    """Same init DISTRIBUTIONS as resnet.py:163-168 (kaiming_normal_, fan_out, relu: N(0, sqrt(2 / (k*k*Cout))); BN weight 1 / bias 0), drawn from a
    numpy Generator so that fixtures are platform-stable."""
    pl, bl = resnet18_layout(in_ch)
    params, bufs = {}, {}
    for name, shape in pl:
        if len(shape) == 4:
            params[name] = torch.from_numpy((rng.standard_normal(shape) * math.sqrt(2.0 / (shape[0] * shape[2] * shape[3]))).astype(np.float32))
        elif name.endswith(".weight"):
            params[name] = torch.ones(shape)
        else:
            params[name] = torch.zeros(shape)
    for name, shape in bl:
        bufs[name] = torch.ones(shape) if name.endswith("running_var") else (torch.zeros((), dtype=torch.int64) if name.endswith("num_batches_tracked")
                                                                           else torch.zeros(shape))
    return params, bufs


class _BF16Conv(torch.autograd.Function):
    """Arithmetic class of the product's ResNet18 / AlexNet path: every contraction (forward, data gradient, weight gradient) multiplies
    BF16-rounded operands (round-to-nearest-even) and accumulates in fp32."""

    @staticmethod
    def forward(ctx, x, w, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.sp = (stride, padding)
        return F.conv2d(x.bfloat16().float(), w.bfloat16().float(), None, stride, padding)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        stride, padding = ctx.sp
        dyr = dy.bfloat16().float()
        dx = torch.nn.grad.conv2d_input(x.shape, w.bfloat16().float(), dyr, stride=stride, padding=padding) if ctx.needs_input_grad[0] else None
        dw = torch.nn.grad.conv2d_weight(x.bfloat16().float(), w.shape, dyr, stride=stride, padding=padding)
        return dx, dw, None, None


def _conv(x: Tensor, w: Tensor, stride: int, padding: int, conv_mode: str) -> Tensor:
    if conv_mode == "bf16":
        return _BF16Conv.apply(x, w, stride, padding)
    return F.conv2d(x, w, None, stride, padding)


def resnet18_forward(p: Dict[str, Tensor], b: Dict[str, Tensor], x: Tensor, train: bool, maxpool: bool = True, conv_mode: str = "fp32",
                     relu=None) -> Dict[str, object]:
    """resnet.py:220-234 (`_forward_impl`) with the 3x3 stride-1 stem of :133-150 (+ MaxPool2d(3, 2, 1) for the imagenet-like datasets) and the
    BasicBlock of :48-64.  Returns {'fmaps': [x_1, x_2, x_3, x_4], 'features': [B, 512]}.
    `relu(name, t)` (tests only) replaces F.relu at the site `name` ('conv1', 'layerN.M.relu1', 'layerN.M.relu2') — used to pin the ReLU routing to
    another implementation's masks, which removes the mask-flip discontinuity from a gradient comparison."""
    rl = (lambda name, t: F.relu(t)) if relu is None else relu
    h = _conv(x, p["conv1.0.weight"], 1, 1, conv_mode)
    h = rl("conv1", _bn(h, p, b, "conv1.1", train))
    if maxpool:
        h = F.max_pool2d(h, 3, 2, 1)
    fmaps = []
    for li, stride in enumerate((1, 2, 2, 2), start=1):
        for k in range(2):
            pre = f"layer{li}.{k}"
            s = stride if k == 0 else 1
            identity = h
            y = _conv(h, p[pre + ".conv1.weight"], s, 1, conv_mode)
            y = rl(pre + ".relu1", _bn(y, p, b, pre + ".bn1", train))
            y = _conv(y, p[pre + ".conv2.weight"], 1, 1, conv_mode)
            y = _bn(y, p, b, pre + ".bn2", train)
            if (pre + ".downsample.0.weight") in p:
                identity = _bn(_conv(h, p[pre + ".downsample.0.weight"], s, 0, conv_mode), p, b, pre + ".downsample.1", train)
            h = rl(pre + ".relu2", y + identity)
        fmaps.append(h)
    feats = torch.flatten(F.adaptive_avg_pool2d(h, (1, 1)), 1)
    return {"fmaps": fmaps, "features": feats}


# ----------------------------------------------------------------------------------------------
# heads and losses
# ----------------------------------------------------------------------------------------------
def linear_head(feat: Tensor, w: Tensor, bias: Optional[Tensor]) -> Tensor:
    """nn.Linear(feat_dim, n_cls) — ewc.py:52-57, icarl.py:24-38, finetune.py:10,19."""
    return F.linear(feat, w, bias)


def cosine_head(feat: Tensor, w: Tensor, sigma: Optional[Tensor]) -> Tensor:
    """CosineLinear.forward, resnet.py:436-441: sigma * normalize(x) @ normalize(W)^T (eps 1e-12)."""
    out = F.linear(F.normalize(feat, p=2, dim=1), F.normalize(w, p=2, dim=1))
    return out if sigma is None else sigma * out


def kd_loss(student: Tensor, teacher: Tensor, T: float = 2.0) -> Tensor:
    """`_KD_loss` of lwf.py:75-78 / icarl.py:198-206: -(1/B) sum softmax(t/T) * log_softmax(s/T).  No T^2 factor."""
    return -(torch.softmax(teacher / T, dim=1) * torch.log_softmax(student / T, dim=1)).sum() / student.shape[0]


def ewc_penalty(named_params: Dict[str, Tensor], ref: Dict[str, Tensor], fisher: Dict[str, Tensor]) -> Tensor:
    """`EWC.compute_ewc`, ewc.py:207-225: sum_n sum(F_n * (p_n[:len(ref_n)] - ref_n)^2) / 2."""
    total = 0.0
    for n, prm in named_params.items():
        if n in fisher:
            total = total + (fisher[n] * (prm[: len(ref[n])] - ref[n]).pow(2)).sum() / 2
    return total


def ewc_loss(logits: Tensor, y: Tensor, task_idx: int, inc_cls: int, lamda: float,
             named_params: Dict[str, Tensor], ref: Dict[str, Tensor], fisher: Dict[str, Tensor]) -> Tensor:
    """`EWC.observe`, ewc.py:82-100."""
    if task_idx == 0:
        return F.cross_entropy(logits, y)
    old = logits.shape[1] - inc_cls
    return F.cross_entropy(logits[:, old:], y - old) + lamda * ewc_penalty(named_params, ref, fisher)


def icarl_loss(cur_logits_all: Tensor, y: Tensor, accu: int, prev: int, old_logits_all: Optional[Tensor]) -> Tuple[Tensor, Tensor]:
    """`ICarl.criterion`, icarl.py:197-221.  `cur_logits_all` = network(x) over the fixed `num_class` head."""
    cur = cur_logits_all[:, :accu]
    loss = F.cross_entropy(cur, y)
    if old_logits_all is not None:
        loss = loss + kd_loss(cur[:, :prev], old_logits_all[:, :prev], 2.0)
    return cur, loss


def lwf_loss(logits: Tensor, y: Tensor, known: int, teacher_logits: Optional[Tensor]) -> Tensor:
    """`LWF.observe`, lwf.py:52-70 (lamda=3 and T=2 are hard-coded at :63-65)."""
    if teacher_logits is None:
        return F.cross_entropy(logits, y)
    return 3 * kd_loss(logits[:, :known], teacher_logits[:, :known], 2.0) + F.cross_entropy(logits[:, known:], y - known)


def lucir_loss(feat: Tensor, ref_feat: Tensor, logits: Tensor, scores_bs: Tensor, y: Tensor, num_old: int,
               cur_lamda: float, K: int, dist: float, lw_mr: float) -> Tensor:
    """`LUCIR.observe` task>0 branch, lucir.py:184-205.
    feat/ref_feat: inputs of the cosine classifiers (student / frozen ref); logits = sigma*scores_bs."""
    B = y.shape[0]
    loss = F.cosine_embedding_loss(feat, ref_feat.detach(), torch.ones(B, device=feat.device)) * cur_lamda
    loss = loss + F.cross_entropy(logits, y)
    gt = scores_bs.gather(1, y.view(-1, 1)).squeeze(1)
    max_novel = scores_bs[:, num_old:].topk(K, dim=1)[0]
    hard = y.lt(num_old)
    n_hard = int(hard.sum())
    if n_hard > 0:
        g = gt[hard].view(-1, 1).repeat(1, K)
        m = max_novel[hard]
        loss = loss + F.margin_ranking_loss(g.view(-1, 1), m.view(-1, 1), torch.ones(n_hard * K, 1), margin=dist) * lw_mr
    return loss


# ----------------------------------------------------------------------------------------------
# optimizers  (torch.optim.SGD / Adam as configured by trainer.py:159-182; restated from their documented update rules)
# ----------------------------------------------------------------------------------------------
def sgd_momentum_step(p: Tensor, g: Tensor, m: Optional[Tensor], lr: float, momentum: float, wd: float) -> Tuple[Tensor, Tensor]:
    """torch.optim.SGD (dampening 0, no nesterov): g' = g + wd*p ; m = g' (first step) or mu*m + g' ; p -= lr*m."""
    g = g.add(p, alpha=wd) if wd != 0 else g
    m = g.clone() if m is None else m.mul(momentum).add_(g)
    return p.add(m, alpha=-lr), m


def adam_step(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1: float, b2: float, eps: float, wd: float):
    """torch.optim.Adam (no amsgrad), L2-style weight decay."""
    g = g + wd * p
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    return p - (lr / bc1) * m / denom, m, v


# ----------------------------------------------------------------------------------------------
# L2P prompt pool selection  (core/model/backbone/prompt.py:369-406)
# ----------------------------------------------------------------------------------------------
def l2p_select(prompt: Tensor, prompt_key: Tensor, cls_features: Tensor, top_k: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Returns (batched_prompt [L, B, top_k*len, D], reduce_sim scalar, major_prompt_id [top_k] int64).
    Batch-wide majority vote (prompt.py:381-390): per-sample top-k ids -> histogram -> top-k most frequent ids,
    shared by every sample."""
    pool = prompt_key.shape[0]
    B = cls_features.shape[0]
    kn = F.normalize(prompt_key, p=2, dim=-1, eps=1e-12)
    qn = F.normalize(cls_features, p=2, dim=-1, eps=1e-12)
    sim = qn @ kn.T
    idx = sim.topk(top_k, dim=1)[1]
    ids, counts = torch.unique(idx, return_counts=True, sorted=True)
    ids = F.pad(ids, (0, pool - len(ids)), "constant", int(ids[0]))
    counts = F.pad(counts, (0, pool - len(counts)), "constant", 0)
    major = ids[counts.topk(top_k)[1]]
    sel = major.unsqueeze(0).repeat(B, 1)
    raw = prompt[:, sel]                                    # [L, B, top_k, len, D]
    batched = raw.reshape(raw.shape[0], raw.shape[1], -1, raw.shape[-1])
    reduce_sim = (kn[sel] * qn.unsqueeze(1)).sum() / B
    return batched, reduce_sim, major


def l2p_majority_ids_numpy(sim: np.ndarray, top_k: int) -> np.ndarray:
    """Integer-only restatement of the id selection above with an explicit tie rule, used to define the
    product's deterministic behaviour: per-sample top-k by (value desc, index asc); histogram over the pool;
    majority top-k by (count desc, id asc).  `torch.topk` on CPU resolves ties the same way for these sizes
    (checked against the reference in tests/test_oracle_golden.py)."""
    B, pool = sim.shape
    hist = np.zeros(pool, dtype=np.int64)
    for b in range(B):
        order = sorted(range(pool), key=lambda j: (-float(sim[b, j]), j))[:top_k]
        for j in order:
            hist[j] += 1
    return l2p_majority_from_hist_numpy(hist, top_k)


def l2p_majority_from_hist_numpy(hist: np.ndarray, top_k: int) -> np.ndarray:
    """The majority step of `l2p_majority_ids_numpy` alone (prompt.py:380-401 after `torch.unique(..., return_counts=True)`): a histogram of
    per-sample top-k picks -> the top_k most frequent prompt ids.  Under data parallelism the histogram is the SUM over the ranks' shards."""
    pool = len(hist)
    present = [j for j in range(pool) if hist[j] > 0]
    # reference pads the id list with ids[0] / count 0 (prompt.py:384-385)
    ids = present + [present[0]] * (pool - len(present))
    cnt = [int(hist[j]) for j in present] + [0] * (pool - len(present))
    order = sorted(range(pool), key=lambda i: (-cnt[i], i))[:top_k]
    return np.array([ids[i] for i in order], dtype=np.int64)


# ----------------------------------------------------------------------------------------------
# ViT-B/16 backbone + L2P  (core/model/backbone/transformer.py:2222-2261 VisionTransformer.forward, :2006-2017 Transformer.forward,
# :1331-1336 ResidualAttentionBlock.forward, :169-197 MultiHeadAttention.forward, :1267-1273 Mlp; core/model/backbone/vit.py:100-121
# ViTZoo.forward (l2p branch); core/model/l2p.py:84-107 observe)
# ----------------------------------------------------------------------------------------------
def vit_layout(depth: int = 12, dim: int = 768, mlp: int = 3072, patch: int = 16, tokens: int = 197) -> List[Tuple[str, Tuple[int, ...]]]:
    """Parameter names (state-dict keys of the reference `VisionTransformer`) and shapes, in registration order."""
    out: List[Tuple[str, Tuple[int, ...]]] = [("cls_token", (1, 1, dim)), ("pos_embed", (1, tokens, dim)),
                                             ("patch_embed.proj.weight", (dim, 3, patch, patch)), ("patch_embed.proj.bias", (dim,))]
    for i in range(depth):
        b = f"transformer.blocks.{i}."
        out += [(b + "attn.qkv.weight", (3 * dim, dim)), (b + "attn.qkv.bias", (3 * dim,)), (b + "attn.proj.weight", (dim, dim)),
                (b + "attn.proj.bias", (dim,)), (b + "ln_1.weight", (dim,)), (b + "ln_1.bias", (dim,)), (b + "mlp.fc1.weight", (mlp, dim)),
                (b + "mlp.fc1.bias", (mlp,)), (b + "mlp.fc2.weight", (dim, mlp)), (b + "mlp.fc2.bias", (dim,)), (b + "ln_2.weight", (dim,)),
                (b + "ln_2.bias", (dim,))]
    out += [("norm.weight", (dim,)), ("norm.bias", (dim,))]
    return out


def vit_init(rng: np.random.Generator, depth: int = 12, dim: int = 768, mlp: int = 3072) -> Dict[str, Tensor]:
    """Synthetic 'pretrained' weights from a numpy Generator (there is no checkpoint offline): matrices N(0, 0.02^2) like
    `_init_weights` (transformer.py:2207-2214) but with non-zero biases / non-unit LayerNorm affine so that every term is exercised."""
    p: Dict[str, Tensor] = {}
    for name, shape in vit_layout(depth, dim, mlp):
        if name.endswith("ln_1.weight") or name.endswith("ln_2.weight") or name == "norm.weight":
            a = 1.0 + 0.1 * rng.standard_normal(shape)
        elif name.endswith(".bias"):
            a = 0.02 * rng.standard_normal(shape)
        elif name == "patch_embed.proj.weight":
            a = rng.uniform(-1.0, 1.0, shape) / math.sqrt(shape[1] * shape[2] * shape[3])
        else:
            a = 0.02 * rng.standard_normal(shape)
        p[name] = torch.from_numpy(a.astype(np.float32))
    return p


def _bf16_round(t: Tensor) -> Tensor:
    """Round-to-nearest-even to BF16 with a straight-through gradient (the CUDA path's GEMM operands are BF16)."""
    return t + (t.detach().bfloat16().float() - t.detach())


def vit_tokens(p: Dict[str, Tensor], x: Tensor, prompts: Optional[Tensor] = None, depth: int = 12, heads: int = 12, gemm_mode: str = "fp32",
               taps: Optional[Dict[str, Tensor]] = None, lora: Optional[Sequence[Dict[str, Tensor]]] = None,
               prefix: Optional[Dict[int, Tuple[Tensor, Tensor]]] = None, attn_inputs: Optional[List[Tensor]] = None, block_eps: float = 1e-5) -> Tensor:
    """`VisionTransformer.forward(prompt_flag='l2p')` up to and including the final LayerNorm: [B, (P +) 197, D].
    `prompts` [B, P, D] are prepended in front of [cls, patches] AFTER the position embedding was added (transformer.py:2240-2251,
    :2010-2014).  gemm_mode 'bf16' rounds every GEMM operand (and the stored qkv / probabilities / GELU output) to BF16 exactly where
    the CUDA path does; accumulation stays fp32.
    `lora[i]` = {'A_k','B_k','A_v','B_v'} (or 'A_q','B_q'): weight-side adapters of block i, W_s + B_s A_s (transformer.py:246-254).
    `prefix[i]` = (pk, pv) [B, P, D]: prefix keys / values concatenated in front of block i's K, V (transformer.py:175-180).
    `block_eps`: LayerNorm eps of the blocks (1e-5 in transformer.py:1289,1315; 1e-6 in vit_inflora.py's timm-style blocks).
    `attn_inputs` (a list) receives ln_1(x) of every block, the matrix InfLoRA's `get_input_matrix` accumulates (transformer.py:242-244)."""
    r = _bf16_round if gemm_mode == "bf16" else (lambda t: t)
    B = x.shape[0]
    D = p["cls_token"].shape[-1]
    w_pe = p["patch_embed.proj.weight"]
    ps = w_pe.shape[-1]
    g = x.shape[-1] // ps
    # conv k=16 s=16 == GEMM over im2col patches (column order c, py, px)
    cols = x.reshape(B, 3, g, ps, g, ps).permute(0, 2, 4, 1, 3, 5).reshape(B, g * g, 3 * ps * ps)
    tok = F.linear(r(cols), r(w_pe.reshape(D, -1)), p["patch_embed.proj.bias"])
    xs = torch.cat([p["cls_token"].expand(B, -1, -1), tok], dim=1) + p["pos_embed"][:, : g * g + 1]
    if prompts is not None:
        xs = torch.cat([prompts, xs], dim=1)
    T = xs.shape[1]
    hd = D // heads
    for i in range(depth):
        b = f"transformer.blocks.{i}."
        h = F.layer_norm(xs, (D,), p[b + "ln_1.weight"], p[b + "ln_1.bias"], block_eps)
        if attn_inputs is not None:
            attn_inputs.append(h.detach())
        w_qkv = p[b + "attn.qkv.weight"]
        if lora is not None:
            slabs = list(w_qkv.chunk(3, dim=0))
            for si, sn in enumerate("qkv"):
                if f"A_{sn}" in lora[i]:
                    slabs[si] = slabs[si] + lora[i][f"B_{sn}"] @ lora[i][f"A_{sn}"]
                if f"delta_{sn}" in lora[i]:                     # a ready-made weight delta (SD-LoRA: scaled sum over the tasks' adapters)
                    slabs[si] = slabs[si] + lora[i][f"delta_{sn}"]
            w_qkv = torch.cat(slabs, dim=0)
        qkv = r(F.linear(r(h), r(w_qkv), p[b + "attn.qkv.bias"]))
        qkv = qkv.reshape(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        if prefix is not None and i in prefix:
            pk, pv = prefix[i]
            k = torch.cat([r(pk).reshape(B, -1, heads, hd).permute(0, 2, 1, 3), k], dim=2)
            v = torch.cat([r(pv).reshape(B, -1, heads, hd).permute(0, 2, 1, 3), v], dim=2)
        attn = r(((q @ k.transpose(-2, -1)) * hd ** -0.5).softmax(dim=-1))
        o = r((attn @ v).transpose(1, 2).reshape(B, T, D))
        xs = xs + F.linear(o, r(p[b + "attn.proj.weight"]), p[b + "attn.proj.bias"])
        h = F.layer_norm(xs, (D,), p[b + "ln_2.weight"], p[b + "ln_2.bias"], block_eps)
        u = r(F.gelu(F.linear(r(h), r(p[b + "mlp.fc1.weight"]), p[b + "mlp.fc1.bias"])))
        xs = xs + F.linear(u, r(p[b + "mlp.fc2.weight"]), p[b + "mlp.fc2.bias"])
        if taps is not None:
            taps[f"block{i}"] = xs.detach()
    return F.layer_norm(xs, (D,), p["norm.weight"], p["norm.bias"], 1e-6)


def l2p_forward(p: Dict[str, Tensor], pool_prompt: Tensor, pool_key: Tensor, x: Tensor, top_k: int, depth: int = 12, heads: int = 12,
                gemm_mode: str = "fp32") -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """`ViTZoo.forward` l2p branch (vit.py:102-119): query pass without prompts (no grad) -> pool selection -> prompted pass; the feature is
    the mean over the prompt positions of the normalised tokens (transformer.py:2255-2258).
    Returns (feat [B, D], reduce_sim, major ids [top_k], cls_features [B, D])."""
    with torch.no_grad():
        cls_features = vit_tokens(p, x, None, depth, heads, gemm_mode)[:, 0]
    batched, reduce_sim, major = l2p_select(pool_prompt, pool_key, cls_features, top_k)
    prompts = batched[0]
    y = vit_tokens(p, x, prompts, depth, heads, gemm_mode)
    return y[:, : prompts.shape[1]].mean(dim=1), reduce_sim, major, cls_features


def l2p_loss(logits: Tensor, y: Tensor, lo: int, hi: int, reduce_sim: Tensor, coeff: float) -> Tuple[Tensor, Tensor]:
    """l2p.py:89-99: logits outside the current task's classes [lo, hi) are set to -inf, CE, minus coeff * reduce_sim."""
    masked = torch.full_like(logits, float("-inf"))
    masked[:, lo:hi] = logits[:, lo:hi]
    return F.cross_entropy(masked, y) - coeff * reduce_sim, masked


# ----------------------------------------------------------------------------------------------
# InfLoRA_OPT on ViT-B/16  (core/model/InfLoRA_opt.py:107-121 SiNet.forward, :175-189 observe; transformer.py:199-274)
# ----------------------------------------------------------------------------------------------
def inflora_logits(p: Dict[str, Tensor], lora: Sequence[Dict[str, Tensor]], head_w: Tensor, head_b: Tensor, x: Tensor, depth: int = 12,
                   heads: int = 12, gemm_mode: str = "fp32") -> Tensor:
    """features = final-LayerNorm cls token of the adapted backbone (vit.py:128-131), logits = classifier_pool[task](features)."""
    feat = vit_tokens(p, x, None, depth, heads, gemm_mode, lora=lora)[:, 0]
    return F.linear(feat, head_w, head_b)


def inflora_input_matrices(p: Dict[str, Tensor], lora: Optional[Sequence[Dict[str, Tensor]]], batches: Sequence[Tensor], depth: int = 12,
                           heads: int = 12) -> List[Tensor]:
    """`update_input_matrix` over a loader (InfLoRA_opt.py:243-245, transformer.py:242-244): per block the running mean over tokens of
    h h^T with h = ln_1(x) (the attention input), in the reference's update order."""
    cur = [torch.zeros(p["cls_token"].shape[-1], p["cls_token"].shape[-1]) for _ in range(depth)]
    n = 0
    with torch.no_grad():
        for x in batches:
            hs: List[Tensor] = []
            vit_tokens(p, x, None, depth, heads, "fp32", lora=lora, attn_inputs=hs)
            m = hs[0].shape[0] * hs[0].shape[1]
            for i, h in enumerate(hs):
                cur[i] = (cur[i] * n + torch.bmm(h.permute(0, 2, 1), h).sum(dim=0)) / (n + m)
            n += m
    return cur


def inflora_init_A(cur_matrix: Tensor, rank: int) -> Tensor:
    """Task-0 adapter basis (InfLoRA_opt.py:248-254): the top-`rank` left singular vectors of the input matrix, scaled by 1/sqrt(3)."""
    U, _, _ = torch.linalg.svd(cur_matrix, full_matrices=False)
    return U[:, :rank].T / math.sqrt(3)


# ----------------------------------------------------------------------------------------------
# InfLoRA (original) on the timm-style ViT of vit_inflora.py  (Attention_LoRA.forward :222-252, ViT_lora_co.forward SiNet.py:20-35, InfLoRA.observe)
# ----------------------------------------------------------------------------------------------
def inflora_orig_logits(p: Dict[str, Tensor], blocks: Sequence[Sequence[Dict[str, Tensor]]], head_w: Tensor, head_b: Tensor, x: Tensor, depth: int = 12,
                        heads: int = 12, gemm_mode: str = "fp32") -> Tensor:
    """blocks[l] = adapters {'A_k','B_k','A_v','B_v'} of block l for tasks 0..t: k and v get x (sum_t B_t A_t)^T added (vit_inflora.py:236-240), i.e. the
    weight-side sum; LayerNorm eps 1e-6 everywhere; feature = cls token of the final norm; logits of the current task's head only."""
    lora = [{"delta_k": sum(ad["B_k"] @ ad["A_k"] for ad in blocks[l]), "delta_v": sum(ad["B_v"] @ ad["A_v"] for ad in blocks[l])} for l in range(depth)]
    feat = vit_tokens(p, x, None, depth, heads, gemm_mode, lora=lora, block_eps=1e-6)[:, 0]
    return F.linear(feat, head_w, head_b)


# ----------------------------------------------------------------------------------------------
# SD-LoRA on ViT-B/16  (core/model/backbone/transformer.py:304-357 MultiHeadAttention_SDLoRA.forward; core/model/sd_lora.py:80-94 observe)
# ----------------------------------------------------------------------------------------------
def sdlora_deltas(adapters: Sequence[Dict[str, Tensor]], mags: Sequence[Tensor], assimilated: Optional[Sequence[float]] = None) -> Dict[str, Tensor]:
    """Weight-side form of one block's adapters: the LAST adapter enters as mag[-1] B A, every earlier one as (mag[i] + assimilated[i]) B_i A_i /
    (|B_i|_F |A_i|_F), skipped when either norm is zero (transformer.py:312-332).  adapters[i] = {'A_q','B_q','A_v','B_v'}."""
    out = {}
    for sn in "qv":
        last = adapters[-1]
        d = mags[-1] * (last[f"B_{sn}"] @ last[f"A_{sn}"])
        for i, ad in enumerate(adapters[:-1]):
            nb, na = torch.norm(ad[f"B_{sn}"]), torch.norm(ad[f"A_{sn}"])
            if nb != 0 and na != 0:
                d = d + (mags[i] + (0.0 if assimilated is None else assimilated[i])) * (ad[f"B_{sn}"] @ ad[f"A_{sn}"]) / (nb * na)
        out[f"delta_{sn}"] = d
    return out


def sdlora_logits(p: Dict[str, Tensor], blocks: Sequence[Sequence[Dict[str, Tensor]]], mags: Sequence[Tensor], head_w: Tensor, head_b: Tensor, x: Tensor,
                  depth: int = 12, heads: int = 12, gemm_mode: str = "fp32") -> Tensor:
    """blocks[l] = the adapters of block l in task order; the magnitudes are shared by all blocks (sd_lora.py:122-125)."""
    lora = [sdlora_deltas(blocks[l], mags) for l in range(depth)]
    feat = vit_tokens(p, x, None, depth, heads, gemm_mode, lora=lora)[:, 0]
    return F.linear(feat, head_w, head_b)


# ----------------------------------------------------------------------------------------------
# DualPrompt on ViT-B/16  (core/model/backbone/prompt.py:231-337 pool, transformer.py:2263-2296 non-l2p branch, vit.py:121-131,
# core/model/dualprompt.py:89-104 observe)
# ----------------------------------------------------------------------------------------------
DUAL_G_LAYERS, DUAL_E_LAYERS = (0, 1), (2, 3, 4)


def dualprompt_prefixes(pool: Dict[str, Tensor], q: Tensor, task_id: int, train: bool):
    """Per-layer prefix (keys, values) and the key-match loss.  pool: 'g_p_{l}' [Lg, D], 'e_p_{l}' [pool, Le, D], 'e_k_{l}' [pool, D].
    Training uses the task id (`task_id_bootstrap`, prompt.py:281-284): loss = sum_l sum_b (1 - cos(q_b, K_l[task])), the query detached;
    inference takes the per-sample top-1 key (prompt.py:290-292).  Returns (prefix dict, loss, selected ids per e-layer)."""
    B = q.shape[0]
    prefix: Dict[int, Tuple[Tensor, Tensor]] = {}
    loss = torch.zeros(())
    ids = {}
    for l in DUAL_E_LAYERS:
        K, pp = pool[f"e_k_{l}"], pool[f"e_p_{l}"]
        cos = F.normalize(q, dim=1).detach() @ F.normalize(K, dim=1).T
        if train:
            loss = loss + (1.0 - cos[:, task_id]).sum()
            P_ = pp[task_id].expand(B, -1, -1)
            ids[l] = torch.full((B,), task_id, dtype=torch.int64)
        else:
            ids[l] = cos.argmax(dim=1)
            P_ = pp[ids[l]]
        i = pp.shape[1] // 2
        prefix[l] = (P_[:, :i], P_[:, i:])
    for l in DUAL_G_LAYERS:
        g = pool[f"g_p_{l}"].expand(B, -1, -1)
        j = g.shape[1] // 2
        prefix[l] = (g[:, :j], g[:, j:])
    return prefix, loss, ids


def dualprompt_forward(p: Dict[str, Tensor], pool: Dict[str, Tensor], x: Tensor, task_id: int, train: bool, depth: int = 12, heads: int = 12,
                       gemm_mode: str = "fp32"):
    """`ViTZoo.forward` prompt branch (vit.py:121-131): no-grad query pass -> cls feature, prefix-tuned pass -> cls feature."""
    with torch.no_grad():
        q = vit_tokens(p, x, None, depth, heads, gemm_mode)[:, 0]
    prefix, loss, ids = dualprompt_prefixes(pool, q, task_id, train)
    feat = vit_tokens(p, x, None, depth, heads, gemm_mode, prefix=prefix)[:, 0]
    return feat, loss, q, ids


def dualprompt_loss(logits: Tensor, y: Tensor, last_out_dim: int, prompt_loss: Tensor) -> Tensor:
    """dualprompt.py:96-100: logits of the previous tasks' classes set to -inf, per-sample CE (weight 1) averaged, plus the pool loss."""
    masked = logits.clone()
    masked[:, :last_out_dim] = float("-inf")
    return prompt_loss + F.cross_entropy(masked, y, reduction="none").mean()


# ----------------------------------------------------------------------------------------------
# CodaPrompt on ViT-B/16  (core/model/backbone/prompt.py:146-214 pool forward; core/model/codaprompt.py:89-104 observe = DualPrompt's)
# ----------------------------------------------------------------------------------------------
CODA_LAYERS = (0, 1, 2, 3, 4)


def codaprompt_prefixes(pool: Dict[str, Tensor], q: Tensor, nk: int):
    """Per-layer prefix (keys, values): P_[b] = sum_k cos(q_b * A_k, K_k) p[k] over the first nk components (task_count stays 0 in the reference:
    `process_task_count` is never called, so s = 0, f = pool_size / n_tasks in training and at inference alike).  pool: 'e_p_{l}' [pool, Lp, D],
    'e_k_{l}', 'e_a_{l}' [pool, D]."""
    prefix: Dict[int, Tuple[Tensor, Tensor]] = {}
    for l in CODA_LAYERS:
        K, A, pp = pool[f"e_k_{l}"][:nk], pool[f"e_a_{l}"][:nk], pool[f"e_p_{l}"][:nk]
        a_q = torch.einsum("bd,kd->bkd", q, A)
        aq_k = torch.einsum("bkd,kd->bk", F.normalize(a_q, dim=2), F.normalize(K, dim=1))
        P_ = torch.einsum("bk,kld->bld", aq_k, pp)
        i = pp.shape[1] // 2
        prefix[l] = (P_[:, :i], P_[:, i:])
    return prefix


def codaprompt_forward(p: Dict[str, Tensor], pool: Dict[str, Tensor], x: Tensor, nk: int, depth: int = 12, heads: int = 12, gemm_mode: str = "fp32"):
    with torch.no_grad():
        q = vit_tokens(p, x, None, depth, heads, gemm_mode)[:, 0]
    feat = vit_tokens(p, x, None, depth, heads, gemm_mode, prefix=codaprompt_prefixes(pool, q, nk))[:, 0]
    return feat, q


# ----------------------------------------------------------------------------------------------
# iCaRL exemplar management
# ----------------------------------------------------------------------------------------------
def herding_select(features: Tensor, targets: Tensor, per_class: int) -> List[int]:
    """Greedy herding of `LinearHerdingBuffer.herding_select` (core/model/buffer/linearherdingbuffer.py:133-163) on features that
    are already L2-normalised and ordered by class.  Returns global row indices."""
    feats = features.clone()
    result: List[int] = []
    for c in np.unique(targets.numpy()):
        ind = np.where(targets.numpy() == c)[0]
        cf = feats[ind]
        mean = cf.mean(0, keepdim=True)
        rs = torch.zeros_like(mean)
        i = 0
        while i < per_class and i < cf.shape[0]:
            cost = (mean - (cf + rs) / (i + 1)).norm(2, 1)
            j = int(cost.argmin())
            result.append(j + int(ind[0]))
            rs += cf[j:j + 1]
            cf[j] = cf[j] + 1e6
            i += 1
    return result


def ncm_classify(feats: Tensor, class_means: Tensor) -> Tensor:
    """`ICarl.NCM_classify` (core/model/icarl.py:122-152)."""
    n, m = feats.shape[0], class_means.shape[0]
    d = torch.pow(feats.unsqueeze(1).expand(n, m, -1) - class_means.unsqueeze(0).expand(n, m, -1), 2).sum(2)
    return torch.argmin(d, dim=1)


# ----------------------------------------------------------------------------------------------
# GPM gradient projection  (core/model/gpm.py:78-81, M = U U^T from gpm.py:124)
# ----------------------------------------------------------------------------------------------
def gpm_project(grad: Tensor, feature_mat: Tensor) -> Tensor:
    sz = grad.shape[0]
    return grad - (grad.view(sz, -1) @ feature_mat).view(grad.shape)


# ----------------------------------------------------------------------------------------------
# AlexNet_TRGP + GPM  (core/model/backbone/alexnet.py:94-156, core/model/gpm.py:22-204)
# ----------------------------------------------------------------------------------------------
ALEXNET_LAYERS = (("conv1", "bn1", (64, 3, 4, 4)), ("conv2", "bn2", (128, 64, 3, 3)), ("conv3", "bn3", (256, 128, 2, 2)), ("fc1", "bn4", (2048, 1024)),
                  ("fc2", "bn5", (2048, 2048)))
ALEXNET_DROP = (0.2, 0.2, 0.5, 0.5, 0.5)          # dropout1 after conv1 / conv2, dropout2 after conv3 / fc1 / fc2 (alexnet.py:127-154)


def alexnet_layout() -> List[Tuple[str, Tuple[int, ...]]]:
    """Parameter (name, shape) list in the reference's registration order (alexnet.py:100-114); no buffers (track_running_stats=False)."""
    out = []
    for w, bn, shape in ALEXNET_LAYERS:
        out += [(w + ".weight", shape), (bn + ".weight", (shape[0],)), (bn + ".bias", (shape[0],))]
    return out


def alexnet_init(rng: np.random.Generator) -> Dict[str, Tensor]:
    """nn.Conv2d / nn.Linear default init DISTRIBUTIONS (kaiming_uniform(a=sqrt(5)) = U(-1/sqrt(fan_in), 1/sqrt(fan_in))), BN weight 1 / bias 0."""
    p = {}
    for name, shape in alexnet_layout():
        if len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            b = 1.0 / math.sqrt(fan_in)
            p[name] = torch.from_numpy(rng.uniform(-b, b, shape).astype(np.float32))
        elif name.endswith(".weight"):
            p[name] = torch.ones(shape)
        else:
            p[name] = torch.zeros(shape)
    return p


class _BF16Linear(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return F.linear(x.bfloat16().float(), w.bfloat16().float())

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        dyr = dy.bfloat16().float()
        return dyr @ w.bfloat16().float(), dyr.t() @ x.bfloat16().float()


def alexnet_forward(p: Dict[str, Tensor], x: Tensor, masks: Optional[Sequence[Tensor]] = None, gemm_mode: str = "fp32", keep: Optional[dict] = None) -> Tensor:
    """`AlexNet_TRGP.forward` (alexnet.py:124-156): [conv -> BN(batch statistics, always: track_running_stats=False) -> ReLU -> dropout -> maxpool(2)] x 3,
    flatten (NCHW order), [linear -> BN1d -> ReLU -> dropout] x 2.  `masks`: five keep-masks (train mode; survivors are scaled by 1/(1-p)); None = eval
    (no dropout).  `keep` (optional dict) receives the INPUT of every TRGP layer, what `compute_input_matrix=True` stashes (alexnet.py:36-37,78-79)."""
    conv = (lambda h, w: _BF16Conv.apply(h, w, 1, 0)) if gemm_mode == "bf16" else (lambda h, w: F.conv2d(h, w))
    lin = (lambda h, w: _BF16Linear.apply(h, w)) if gemm_mode == "bf16" else (lambda h, w: F.linear(h, w))
    h = x
    for i, (wn, bn, shape) in enumerate(ALEXNET_LAYERS):
        if i == 3:
            h = h.reshape(h.shape[0], -1)
        if keep is not None:
            keep[wn] = h.detach()
        h = conv(h, p[wn + ".weight"]) if i < 3 else lin(h, p[wn + ".weight"])
        h = F.relu(F.batch_norm(h, None, None, p[bn + ".weight"], p[bn + ".bias"], True, 0.1, 1e-5))
        if masks is not None:
            h = h * masks[i].to(h.dtype) / (1.0 - ALEXNET_DROP[i])
        if i < 3:
            h = F.max_pool2d(h, 2)
    return h


def gpm_representation_matrices(inputs: Dict[str, Tensor]) -> List[np.ndarray]:
    """gpm.py:144-168: im2col of the first 24 / 100 / 100 stashed conv inputs (the reference's Python triple loop, restated with unfold: column
    (n, i, j) = the (c, kh, kw)-flattened patch at (i, j)), and the transposed 125-sample inputs of the two linear layers."""
    mats = []
    for (wn, _, shape), bsz in zip(ALEXNET_LAYERS[:3], (24, 100, 100)):
        act = inputs[wn][:bsz].double()
        cols = F.unfold(act, shape[2])                                   # [bsz][C*k*k][s*s], positions row-major (i, j)
        mats.append(cols.permute(1, 0, 2).reshape(cols.shape[1], -1).numpy())
    for wn, _, _ in ALEXNET_LAYERS[3:]:
        mats.append(inputs[wn].double().numpy().T)
    return mats


def gpm_update_bases(feature_list: List[np.ndarray], mats: Sequence[np.ndarray], task_idx: int) -> List[np.ndarray]:
    """gpm.py:170-204: grow the per-layer bases so that they hold `threshold = 0.97 + 0.003 * task_idx` of each representation's energy."""
    threshold = 0.97 + task_idx * 0.003
    if task_idx == 0:
        out = []
        for activation in mats:
            U, S, _ = np.linalg.svd(activation, full_matrices=False)
            ratio = (S ** 2) / (S ** 2).sum()
            r = int(np.sum(np.cumsum(ratio) < threshold))
            out.append(U[:, :r])
        return out
    out = list(feature_list)
    for i, activation in enumerate(mats):
        _, S, _ = np.linalg.svd(activation, full_matrices=False)
        total = (S ** 2).sum()
        act_hat = activation - out[i] @ out[i].T @ activation
        U, S, _ = np.linalg.svd(act_hat, full_matrices=False)
        hat = (S ** 2).sum()
        ratio = (S ** 2) / total
        accumulated = (total - hat) / total
        if accumulated >= threshold:
            continue
        r = int(np.sum(np.cumsum(ratio) + accumulated < threshold)) + 1
        Ui = np.hstack((out[i], U[:, :r]))
        out[i] = Ui[:, :min(Ui.shape[0], Ui.shape[1])]
    return out


class GPMOracle:
    """`GPM` (gpm.py:43-206) on AlexNet_TRGP: per-task bias-free heads, CE on the current head, gradient projection of the five TRGP layers for
    task > 0, BN affine frozen after task 0, plain SGD."""

    def __init__(self, p: Dict[str, Tensor], heads: Sequence[Tensor], init_cls: int, inc_cls: int, lr: float = 0.01, gemm_mode: str = "fp32"):
        self.p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        self.heads = [h.clone().requires_grad_(True) for h in heads]
        self.init_cls, self.inc_cls, self.lr, self.gemm_mode = init_cls, inc_cls, lr, gemm_mode
        self.cur_task, self.known = 0, 0
        self.feature_list: List[np.ndarray] = []
        self.feature_mat: List[Tensor] = []

    def before_task(self, task_idx: int):
        self.cur_task = task_idx
        if task_idx == 1:
            self.known += self.init_cls
        elif task_idx > 1:
            self.known += self.inc_cls
        if task_idx > 0:
            self.feature_mat = [torch.tensor(f @ f.T, dtype=torch.float32) for f in self.feature_list]

    def trainable(self) -> Dict[str, Tensor]:
        d = {k: v for k, v in self.p.items() if not (self.cur_task > 0 and "bn" in k)}
        d[f"classifiers.{self.cur_task}.weight"] = self.heads[self.cur_task]
        return d

    def step(self, x: Tensor, y: Tensor, masks=None, apply_update: bool = True):
        feat = alexnet_forward(self.p, x, masks, self.gemm_mode)
        logits = F.linear(feat, self.heads[self.cur_task])
        loss = F.cross_entropy(logits, y - self.known)
        tr = self.trainable()
        grads = dict(zip(tr.keys(), torch.autograd.grad(loss, list(tr.values()))))
        if self.cur_task > 0:
            for i, (wn, _, _) in enumerate(ALEXNET_LAYERS):
                grads[wn + ".weight"] = gpm_project(grads[wn + ".weight"], self.feature_mat[i])
        pred = logits.argmax(1)
        if apply_update:
            with torch.no_grad():
                for k, v in tr.items():
                    v.sub_(self.lr * grads[k])
        return pred, float((pred == (y - self.known)).sum()) / x.shape[0], loss.detach(), grads

    def after_task(self, x125: Tensor):
        keep = {}
        with torch.no_grad():
            alexnet_forward({k: v.detach() for k, v in self.p.items()}, x125, None, self.gemm_mode, keep=keep)
        self.feature_list = gpm_update_bases(self.feature_list, gpm_representation_matrices(keep), self.cur_task)


# ----------------------------------------------------------------------------------------------
# InfLoRA_OPT weight-side LoRA (core/model/backbone/transformer.py:246-254)
# ----------------------------------------------------------------------------------------------
def lora_merge_qkv(qkv_w: Tensor, A_k: Tensor, B_k: Tensor, A_v: Tensor, B_v: Tensor) -> Tensor:
    """W' = cat(W_q, W_k + B_k A_k, W_v + B_v A_v).  qkv_w [3D, D]; A [r, D]; B [D, r]."""
    D = qkv_w.shape[1]
    wq, wk, wv = qkv_w[:D], qkv_w[D:2 * D], qkv_w[2 * D:]
    return torch.cat([wq, wk + B_k @ A_k, wv + B_v @ A_v], dim=0)


# ----------------------------------------------------------------------------------------------
# Method-level steppers used by parity tests, the CPU baseline and `bench.py --impl reference`
# ----------------------------------------------------------------------------------------------
class ResNetMethodOracle:
    """EWC / iCaRL / LwF / Finetune on cifar_resnet32 with the reference's step order
    (observe -> zero_grad -> backward -> SGD step; trainer.py:601-606).

    State: `p` (backbone params + 'classifier.weight'/'classifier.bias'), `b` (BN buffers), optional teacher copy,
    EWC `ref`/`fisher` dicts (names follow ewc.py `self.network.named_parameters()`: 'backbone.<n>', 'classifier.<n>')."""

    def __init__(self, method: str, p: Dict[str, Tensor], b: Dict[str, Tensor], fc_w: Tensor, fc_b: Tensor, *,
                 init_cls: int, inc_cls: int, lamda: float = 1000.0, lr: float = 0.1, momentum: float = 0.9, wd: float = 5e-4,
                 depth: int = 32, conv_mode: str = "fp32", arch: str = "cifar_resnet", maxpool: bool = True):
        assert method in ("finetune", "ewc", "icarl", "lwf") and arch in ("cifar_resnet", "resnet18")
        self.method, self.depth, self.conv_mode, self.arch, self.maxpool = method, depth, conv_mode, arch, maxpool
        self.p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        self.b = {k: v.clone() for k, v in b.items()}
        self.fc_w = fc_w.clone().requires_grad_(True)
        self.fc_b = fc_b.clone().requires_grad_(True)
        self.init_cls, self.inc_cls, self.lamda = init_cls, inc_cls, lamda
        self.lr, self.mu, self.wd = lr, momentum, wd
        self.task_idx = 0
        self.mom: Dict[str, Optional[Tensor]] = {}
        self.teacher = None            # (p, b, fc_w, fc_b) frozen copies
        self.ref: Dict[str, Tensor] = {}
        self.fisher: Dict[str, Tensor] = {}
        self.prev_cls = 0
        self.accu_cls = init_cls

    # -- helpers -------------------------------------------------------------------------------
    def named(self) -> Dict[str, Tensor]:
        d = {"backbone." + k: v for k, v in self.p.items()}
        d["classifier.weight"] = self.fc_w
        d["classifier.bias"] = self.fc_b
        return d

    def backbone(self, p, b, x: Tensor, train: bool):
        if self.arch == "resnet18":
            return resnet18_forward(p, b, x, train, self.maxpool, self.conv_mode)
        return cifar_resnet_forward(p, b, x, train, self.depth, self.conv_mode)

    def logits(self, x: Tensor, train: bool) -> Tensor:
        return linear_head(self.backbone(self.p, self.b, x, train)["features"], self.fc_w, self.fc_b)

    def teacher_logits(self, x: Tensor) -> Tensor:
        tp, tb, tw, tbias = self.teacher
        with torch.no_grad():
            return linear_head(self.backbone(tp, tb, x, False)["features"], tw, tbias)

    def snapshot_teacher(self):
        self.teacher = ({k: v.detach().clone() for k, v in self.p.items()}, {k: v.clone() for k, v in self.b.items()},
                        self.fc_w.detach().clone(), self.fc_b.detach().clone())

    def grow_head(self, new_w: Tensor, new_b: Tensor):
        """ewc.py:71-80 / lwf.py:28-42: new Linear whose first rows are the old head."""
        n_old = self.fc_w.shape[0]
        w, bias = new_w.clone(), new_b.clone()
        w[:n_old] = self.fc_w.detach()
        bias[:n_old] = self.fc_b.detach()
        self.fc_w, self.fc_b = w.requires_grad_(True), bias.requires_grad_(True)

    def reset_optimizer(self):
        self.mom = {}       # trainer.py:294 rebuilds the optimizer every task

    # -- one training step ---------------------------------------------------------------------
    def loss(self, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor]:
        lg = self.logits(x, True)
        if self.method == "finetune":
            return lg, F.cross_entropy(lg, y)
        if self.method == "ewc":
            return lg, ewc_loss(lg, y, self.task_idx, self.inc_cls, self.lamda, self.named(), self.ref, self.fisher)
        if self.method == "icarl":
            old = self.teacher_logits(x) if self.teacher is not None else None
            cur, l = icarl_loss(lg, y, self.accu_cls, self.prev_cls, old)
            return cur, l
        old = self.teacher_logits(x) if self.teacher is not None else None
        return lg, lwf_loss(lg, y, self.prev_cls, old)

    def step(self, x: Tensor, y: Tensor, apply_update: bool = True):
        """Returns (pred, acc, loss, grads dict)."""
        lg, loss = self.loss(x, y)
        named = self.named()
        grads = torch.autograd.grad(loss, list(named.values()), allow_unused=True)
        gd = {n: (g if g is not None else torch.zeros_like(v)) for (n, v), g in zip(named.items(), grads)}
        # BASELINE config C5 ("LwF + GPM"): every conv gradient loses its component inside span(U_l), exactly gpm.py:78-81
        for n, M in (getattr(self, "proj", None) or {}).items():
            gd[n] = gpm_project(gd[n], M)
        pred = lg.argmax(dim=1)
        acc = float((pred == y).sum()) / x.shape[0]
        if apply_update:
            with torch.no_grad():
                for n, v in named.items():
                    newp, m = sgd_momentum_step(v.detach(), gd[n], self.mom.get(n), self.lr, self.mu, self.wd)
                    self.mom[n] = m
                    v.copy_(newp)
        return pred, acc, loss.detach(), gd

    # -- EWC task boundary (ewc.py:110-133, 147-205) ---------------------------------------------
    def ewc_after_task(self, batches: Sequence[Tuple[Tensor, Tensor]], loader_batch_size: int):
        named = self.named()
        self.ref = {n: v.detach().clone() for n, v in named.items()}
        new_f = {n: torch.zeros_like(v) for n, v in named.items()}
        for x, y in batches:
            lg = self.logits(x, True)                       # train() mode: BN stats keep moving (ewc.py:182)
            l = F.cross_entropy(lg, y)                      # over ALL logits (ewc.py:192-193)
            gs = torch.autograd.grad(l, list(self.named().values()))
            for (n, _), g in zip(named.items(), gs):
                new_f[n] += g.pow(2) * len(y)               # ewc.py:173
        n_samples = loader_batch_size * len(batches)        # ewc.py:202 (over-counts a ragged last batch)
        new_f = {n: f / n_samples for n, f in new_f.items()}
        alpha = 1 - self.inc_cls / self.fc_w.shape[0]       # ewc.py:129
        for n, old in self.fisher.items():
            new_f[n][: len(old)] = alpha * old + (1 - alpha) * new_f[n][: len(old)]
        self.fisher = new_f
