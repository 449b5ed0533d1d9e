"""`cifar_resnet32` backbone (mirrors core/model/backbone/resnet.py:324-412 and the factory at :760-763).

The nn.Module keeps the reference's parameter / buffer names (`conv_1_3x3.weight`, `stage_1.0.bn_a.running_mean`, ...), so
`state_dict()` / `load_state_dict()` interoperate with reference checkpoints, but every tensor is a view into the flat
arenas of a `ResNetEngine`, and `forward` runs the hand-written CUDA kernels (autograd-aware through `_BackboneFn`).
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from ...engine import ResNetEngine


class _BackboneFn(torch.autograd.Function):
    """features = backbone(x) with a hand-written backward.  Inputs: x, then every backbone parameter (so that autograd
    routes the gradients); the arithmetic reads the arenas directly."""

    @staticmethod
    def forward(ctx, module, x, *params):
        eng: ResNetEngine = module.engine
        train = module.training
        eng.forward(x, train=train, update_running=train)
        eng.pool_forward(x.shape[0])
        if train:
            module._bump_num_batches_tracked()
        B = x.shape[0]
        ctx.module, ctx.x, ctx.train = module, x, train
        feat = eng.features(B).clone()
        fm = [f.clone() for f in eng.fmaps(B)]
        ctx.mark_non_differentiable(*fm)
        return (feat, *fm)

    @staticmethod
    def backward(ctx, gfeat, *_unused):
        module, x = ctx.module, ctx.x
        if not ctx.train:
            raise RuntimeError("backward through the backbone is only supported in train() mode (batch-statistics BN)")
        eng: ResNetEngine = module.engine
        eng.pool_backward(gfeat.contiguous())
        eng.backward(x)
        eng.autograd_grads = eng.grads.clone()      # see _ArenaLoss.backward: autograd never aliases the live arena
        grads = tuple(eng.param_view(n, eng.autograd_grads) for n, _ in eng.layout)
        return (None, None, *grads)


class CifarResNet(nn.Module):
    def __init__(self, depth: int = 32, channels: int = 3, device=None, max_batch: int = 128, num_class_cap: int = 100, precision: str = "fp32",
                 style: str = "cifar", last_relu: bool = True):
        super().__init__()
        self.engine = ResNetEngine(depth=depth, max_batch=max_batch, num_class_cap=num_class_cap, device=device, in_ch=channels, style=style,
                                   last_relu=last_relu)
        self.engine.set_precision(precision)
        self.out_dim = 64
        self.num_batches_pending = 0
        self._nbt = []
        eng = self.engine
        # same init distributions and the same RNG draw order as resnet.py:345-352
        for name, shape in eng.layout:
            view = eng.param_view(name)
            if len(shape) == 4:
                n = shape[2] * shape[3] * shape[0]
                view.copy_(torch.empty(shape).normal_(0, math.sqrt(2.0 / n)))
            elif name.endswith(".weight"):
                view.fill_(1.0)
            else:
                view.zero_()
            self._register(name, nn.Parameter(view))
        for bn in eng.bn_names:
            m, v = eng.running_views(bn)
            self._register_buffer(bn + ".running_mean", m)
            self._register_buffer(bn + ".running_var", v)
            self._register_buffer(bn + ".num_batches_tracked", torch.zeros((), dtype=torch.long))

    # registration under the reference's dotted names: a tree of bare container modules (`stage_1` -> `0` -> `conv_a` ...) whose leaves hold the arena
    # views, so that named_parameters() / state_dict() / load_state_dict() give the reference's keys at this level AND through any parent module
    # (nn.Module's own recursion does the work: no renaming overrides that a parent would bypass)
    def _leaf(self, name: str):
        *path, leaf = name.split(".")
        mod = self
        for part in path:
            nxt = mod._modules.get(part)
            if nxt is None:
                nxt = _Node()
                mod.add_module(part, nxt)
            mod = nxt
        return mod, leaf

    def _register(self, name, p):
        mod, leaf = self._leaf(name)
        mod.register_parameter(leaf, p)

    def _register_buffer(self, name, b):
        mod, leaf = self._leaf(name)
        mod.register_buffer(leaf, b)
        if leaf == "num_batches_tracked":
            self._nbt.append(b)

    def state_dict(self, *args, **kwargs):
        self._flush_num_batches()
        return super().state_dict(*args, **kwargs)

    def _apply(self, fn, recurse=True):
        # parameters are views into the engine arenas: moving/casting them would silently detach them from the kernels
        probe = fn(torch.empty(0, device=self.engine.device))
        if probe.device != self.engine.device or probe.dtype != torch.float32:
            raise RuntimeError("libcontinual_b200 backbones live on the device they were built on, in fp32")
        return self

    def _bump_num_batches_tracked(self, k: int = 1):
        self.num_batches_pending += k

    def _flush_num_batches(self):
        if self.num_batches_pending:
            for t in self._nbt:
                t.add_(self.num_batches_pending)
            self.num_batches_pending = 0

    # reference interface -------------------------------------------------------------------------------------------------
    def forward(self, x):
        x = x.to(self.engine.device, torch.float32).contiguous()
        if torch.is_grad_enabled() and self.training:
            outs = _BackboneFn.apply(self, x, *[p for _, p in self.named_parameters()])
        else:
            with torch.no_grad():
                outs = _BackboneFn.forward(_NoCtx(), self, x)
        return {"fmaps": list(outs[1:]), "features": outs[0]}

    def feature(self, x):
        return self.forward(x)["features"]


class ResNet18(CifarResNet):
    """torchvision-style ResNet18 of the reference (`ResNet(BasicBlock, [2, 2, 2, 2])`, resnet.py:110-246, factory :259-267) on `ResNet18Engine`.
    Parameter / buffer names follow the reference (`conv1.0.weight`, `layer2.0.downsample.1.running_mean`, ...); the initial weights consume the torch
    RNG exactly as the reference constructor does (constructor draws of every nn.Conv2d, then kaiming_normal_(fan_out) in `modules()` order, then the
    unused `fc = nn.Linear(512, 20)` of resnet.py:191)."""

    def __init__(self, device=None, max_batch: int = 256, num_class_cap: int = 200, img: int = 64, maxpool: bool = True):
        nn.Module.__init__(self)
        from ...nn_engine import ResNet18Engine
        self.engine = ResNet18Engine(max_batch=max_batch, num_class_cap=num_class_cap, device=device, img=img, maxpool=maxpool)
        self.out_dim = 512
        self.num_batches_pending = 0
        self._nbt = []
        eng = self.engine
        shapes = dict(eng.layout)
        draws = {}
        order = []
        # construction order: stem conv; per layer `_make_layer` builds the downsample conv BEFORE the block's conv1 / conv2 (resnet.py:203-212)
        order.append("conv1.0.weight")
        for li in range(1, 5):
            for b in range(2):
                pre = f"layer{li}.{b}"
                if pre + ".downsample.0.weight" in shapes:
                    order.append(pre + ".downsample.0.weight")
                order += [pre + ".conv1.weight", pre + ".conv2.weight"]
        for name in order:
            co, ci, kh, kw = shapes[name]
            draws[name] = nn.Conv2d(ci, co, kh, bias=False).weight.data          # the constructor's kaiming_uniform draw (discarded below)
        for name, shape in eng.layout:                                           # `for m in self.modules()` (resnet.py:163-168): registration order
            if len(shape) == 4:
                nn.init.kaiming_normal_(draws[name], mode="fan_out", nonlinearity="relu")
        fc = nn.Linear(512, 20)
        for name, shape in eng.layout:
            view = eng.param_view(name)
            if len(shape) == 4:
                view.copy_(draws[name])
            elif name.endswith(".weight"):
                view.fill_(1.0)
            else:
                view.zero_()
            self._register(name, nn.Parameter(view))
        for bn in eng.bn_names:
            m, v = eng.running_views(bn)
            self._register_buffer(bn + ".running_mean", m)
            self._register_buffer(bn + ".running_var", v)
            self._register_buffer(bn + ".num_batches_tracked", torch.zeros((), dtype=torch.long))
        # the reference's unused head (never in the forward: no gradient, never updated); kept so that state_dicts interoperate
        self._register("fc.weight", nn.Parameter(fc.weight.data.to(eng.device)))
        self._register("fc.bias", nn.Parameter(fc.bias.data.to(eng.device)))


def resnet18(pretrained: bool = False, progress: bool = True, **kwargs):
    """Factory named in the YAML recipes (`backbone.name: resnet18`, resnet.py:259-267).  `args` selects the stem like resnet.py:133-150:
    'cifar' / '5-datasets' -> 3x3 stride-1 conv; '*imagenet*' with init_cls_num != inc_cls_num -> the same conv + MaxPool(3, 2, 1).  The 7x7 stride-2
    ImageNet stem (init_cls_num == inc_cls_num) is not built."""
    if pretrained:
        raise NotImplementedError
    args = kwargs.get("args") or {}
    ds = str(args.get("dataset", "tiny-imagenet"))
    if "cifar" in ds or "5-datasets" in ds:
        maxpool, img = False, kwargs.get("img", 32)
    elif "imagenet" in ds:
        if args.get("init_cls_num") is not None and args.get("init_cls_num") == args.get("inc_cls_num"):
            raise NotImplementedError("the 7x7 stride-2 stem (resnet.py:136-142) is not built; BASELINE config C5 has init_cls_num != inc_cls_num")
        maxpool, img = True, kwargs.get("img", 64)
    else:
        raise ValueError(f"resnet18: unknown dataset {ds!r} (resnet.py:133-150 knows cifar / 5-datasets / *imagenet*)")
    return ResNet18(device=kwargs.get("device"), max_batch=kwargs.get("max_batch", 256), num_class_cap=kwargs.get("num_classes", 200), img=img,
                    maxpool=maxpool)


class _Node(nn.Module):
    """Bare container (one path component of a reference parameter name).  Loading copies IN PLACE (nn.Module's default for
    parameters and buffers), so the arenas keep owning the storage."""


class _NoCtx:
    def mark_non_differentiable(self, *a):
        pass


def cifar_resnet32(pretrained: bool = False, **kwargs):
    """Factory named in the YAML recipes (`backbone.name: cifar_resnet32`, resnet.py:760-763).  Reference kwargs
    (`num_classes`, `args`) are accepted and ignored exactly as the reference ignores them."""
    return CifarResNet(32, device=kwargs.get("device"), max_batch=kwargs.get("max_batch", 128), num_class_cap=kwargs.get("num_classes", 100),
                       precision=kwargs.get("precision", "fp32"))


def cifar_resnet20(pretrained: bool = False, **kwargs):
    return CifarResNet(20, device=kwargs.get("device"), max_batch=kwargs.get("max_batch", 128), num_class_cap=kwargs.get("num_classes", 100),
                       precision=kwargs.get("precision", "fp32"))


def resnet32_V2(pretrained: bool = False, **kwargs):
    """LUCIR's backbone (`modified_ResNet(modified_BasicBlock, [5, 5, 5])`, resnet.py:472-547, factory :770-774): the same topology
    as cifar_resnet32 with `layerN.M.conv1/bn1/conv2/bn2` names and NO ReLU after the last residual block."""
    return CifarResNet(32, device=kwargs.get("device"), max_batch=kwargs.get("max_batch", 128), num_class_cap=kwargs.get("num_classes", 100),
                       precision=kwargs.get("precision", "fp32"), style="lucir", last_relu=False)
