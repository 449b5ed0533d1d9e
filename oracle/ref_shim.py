"""TEST INFRASTRUCTURE ONLY — never imported by the product path.

Makes the upstream reference (`/root/reference`, RL-VIG/LibContinual @ a399ed4) importable in THIS
container so that `oracle/make_golden.py` can (a) validate the oracle restatement in `oracle/port.py`
against the real reference classes and (b) dump golden vectors into `tests/golden/`.

The reference needs a handful of un-vendored third-party modules at import time (timm, continuum,
diffdist, ftfy, easydict, matplotlib).  They are replaced by the minimal
stand-ins below.  Only `timm.models.vision_transformer.PatchEmbed`, `timm.models.layers.DropPath`,
`trunc_normal_` and `Mlp` carry arithmetic; they follow timm 0.6.7 (the version pinned by
`/root/reference/requirements.txt:9`).

`/root/reference` does not exist on the GPU box, so nothing here may be touched by `-m gpu` tests,
`__graft_entry__.smoke()` or `bench.py`.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LC_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "core"))


def _mod(name: str, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so that "import a.b" works
    sys.modules[name] = m
    return m


def _placeholder(name):
    def _f(*a, **k):
        raise RuntimeError(f"stub for un-vendored dependency called: {name}")
    _f.__name__ = name
    return _f


def install_stubs() -> None:
    import torch
    import torch.nn as nn

    if "timm" in sys.modules and getattr(sys.modules["timm"], "_lc_stub", False):
        return

    class PatchEmbed(nn.Module):
        """timm 0.6.7 PatchEmbed: Conv2d(in, embed, k=patch, s=patch) -> flatten(2) -> transpose(1, 2)."""

        def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=768, norm_layer=None, flatten=True):
            super().__init__()
            self.img_size = (img_size, img_size)
            self.patch_size = (patch_size, patch_size)
            self.grid_size = (img_size // patch_size, img_size // patch_size)
            self.num_patches = self.grid_size[0] * self.grid_size[1]
            self.flatten = flatten
            self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size)
            self.norm = norm_layer(embed_dim) if norm_layer else nn.Identity()

        def forward(self, x):
            x = self.proj(x)
            if self.flatten:
                x = x.flatten(2).transpose(1, 2)
            return self.norm(x)

    class DropPath(nn.Module):
        def __init__(self, drop_prob=0.0, scale_by_keep=True):
            super().__init__()
            self.drop_prob = drop_prob

        def forward(self, x):
            assert self.drop_prob == 0.0 or not self.training, "DropPath>0 not modelled by the stub"
            return x

    class Mlp(nn.Module):
        def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0.0):
            super().__init__()
            out_features = out_features or in_features
            hidden_features = hidden_features or in_features
            self.fc1 = nn.Linear(in_features, hidden_features)
            self.act = act_layer()
            self.drop1 = nn.Dropout(drop)
            self.fc2 = nn.Linear(hidden_features, out_features)
            self.drop2 = nn.Dropout(drop)

        def forward(self, x):
            return self.drop2(self.fc2(self.drop1(self.act(self.fc1(x)))))

    def to_2tuple(x):
        return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

    def register_model(fn):
        return fn

    def named_apply(fn, module, name="", depth_first=True, include_root=False):
        """timm.models.helpers.named_apply (0.6.7 semantics): fn(module=, name=) over the module tree, children before parents by default;
        used by vit_inflora.VisionTransformer.init_weights (vit_inflora.py:428-435)."""
        if not depth_first and include_root:
            fn(module=module, name=name)
        for cname, child in module.named_children():
            named_apply(fn, child, ".".join((name, cname)) if name else cname, depth_first, True)
        if depth_first and include_root:
            fn(module=module, name=name)
        return module

    def _cfg(url="", **kwargs):
        return dict(url=url, **kwargs)

    mean_std = dict(
        IMAGENET_DEFAULT_MEAN=(0.485, 0.456, 0.406), IMAGENET_DEFAULT_STD=(0.229, 0.224, 0.225),
        IMAGENET_INCEPTION_MEAN=(0.5, 0.5, 0.5), IMAGENET_INCEPTION_STD=(0.5, 0.5, 0.5),
    )
    timm = _mod("timm", create_model=_placeholder("timm.create_model"), _lc_stub=True)
    timm.data = _mod("timm.data", **mean_std)
    timm.models = _mod("timm.models", create_model=_placeholder("timm.models.create_model"))
    timm.models.vision_transformer = _mod(
        "timm.models.vision_transformer", _cfg=_cfg, PatchEmbed=PatchEmbed,
        VisionTransformer=type("VisionTransformer", (nn.Module,), {}),
        checkpoint_filter_fn=_placeholder("checkpoint_filter_fn"), default_cfgs={},
    )
    timm.models.layers = _mod(
        "timm.models.layers", trunc_normal_=torch.nn.init.trunc_normal_, DropPath=DropPath, PatchEmbed=PatchEmbed,
        Mlp=Mlp, lecun_normal_=_placeholder("lecun_normal_"), _assert=lambda c, m="": None, to_2tuple=to_2tuple,
    )
    timm.models.layers.helpers = _mod("timm.models.layers.helpers", to_2tuple=to_2tuple)
    timm.models.helpers = _mod(
        "timm.models.helpers", named_apply=named_apply, adapt_input_conv=_placeholder("adapt_input_conv"),
        build_model_with_cfg=_placeholder("build_model_with_cfg"),
        resolve_pretrained_cfg=_placeholder("resolve_pretrained_cfg"), checkpoint_seq=_placeholder("checkpoint_seq"),
    )
    timm.models.registry = _mod("timm.models.registry", register_model=register_model)

    cont = _mod("continuum", ClassIncremental=_placeholder("ClassIncremental"))
    cont.datasets = _mod("continuum.datasets", TinyImageNet200=_placeholder("TinyImageNet200"))
    dd = _mod("diffdist")
    dd.functional = _mod("diffdist.functional")
    _mod("ftfy")
    _mod("easydict", EasyDict=dict)
    try:
        import matplotlib  # noqa: F401
    except Exception:
        mpl = _mod("matplotlib")
        mpl.pyplot = _mod("matplotlib.pyplot")


_core = None


def import_reference():
    """Returns the reference's `core` package (imported from REFERENCE_ROOT with the stubs installed)."""
    global _core
    if _core is not None:
        return _core
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True  # the reference tree is read-only
    install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import core  # noqa: E402

    _core = core
    return core
