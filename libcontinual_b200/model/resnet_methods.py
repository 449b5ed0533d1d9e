"""Finetune / EWC / iCaRL / LwF on the CIFAR ResNet — host-side mirror of the reference plugin classes
(core/model/finetune.py, ewc.py, icarl.py, lwf.py): same constructor kwargs, `observe / inference / before_task /
after_task / get_parameters` with the same return tuples, same RNG draw order for freshly initialised heads.

`observe` issues one fixed kernel sequence (backbone fwd -> head -> CE/KD loss -> head bwd -> backbone bwd -> EWC penalty)
that leaves every gradient in the flat gradient arena, and returns a `loss` whose `.backward()` hands those gradients to
autograd unchanged, so the reference `Trainer._train` order (observe -> zero_grad -> backward -> step, trainer.py:601-606)
runs unmodified on top of it.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .._lib import host_acc

from ..engine import ResNetEngine, TeacherState
from .backbone.resnet import CifarResNet


class _ArenaLoss(torch.autograd.Function):
    """loss tensor whose backward hands the gradients the fused step already produced to the parameters.

    Default: the autograd graph has ONE input (a one-element anchor tensor), and `backward` assigns `p.grad` itself — 101 cached views of a persistent
    flat copy of the gradient arena scaled by the incoming gradient (one kernel).  Returning 101 tensors through the autograd engine instead costs
    ~0.5 ms of host time per step (a view object + an AccumulateGrad node each), a third of a whole 128-image ResNet32 step.  When a parameter already
    holds a gradient (no `zero_grad()` since the last backward: gradient accumulation) the new gradient is added to it, as autograd would.
    The flat copy is reused by the next backward that follows a `zero_grad()`: a `p.grad` tensor kept across steps sees the new values (clone it to keep it).
    `LC_B200_AUTOGRAD_VIEWS=1` restores the engine-mediated hand-off (needed only for `torch.autograd.grad(loss, params)`-style callers)."""

    @staticmethod
    def forward(ctx, owner, loss_value, *inputs):
        ctx.owner = owner
        ctx.direct = len(inputs) == 1 and inputs[0] is owner.__dict__.get("_anchor")
        return loss_value.clone()

    @staticmethod
    def backward(ctx, g):
        owner = ctx.owner
        eng = owner.engine
        if not ctx.direct:
            # autograd gets its own copy (one 1.9 MB kernel): p.grad must never alias the live arena, which the next fused step
            # overwrites, or gradient accumulation / in-place clipping on p.grad would corrupt either side
            eng.autograd_grads = eng.grads * g
            return (None, None, *owner._grad_views_fast(eng.autograd_grads))
        params = owner._params_cached()
        if all(p.grad is None for p in params):          # the Trainer's order: zero_grad() precedes backward() (trainer.py:601-604)
            buf, views = owner._grad_buffer()
            torch.mul(eng.grads, g, out=buf)
            for p, v in zip(params, views):
                if p.requires_grad:
                    p.grad = v
            eng.autograd_grads = buf
        else:
            fresh = eng.grads * g
            for p, v in zip(params, owner._grad_views_fast(fresh)):
                if p.requires_grad:
                    p.grad = v if p.grad is None else p.grad + v
            eng.autograd_grads = None                    # p.grad are no longer views of one flat buffer: the optimizer gathers them
        return (None, None, None)


class _Head(nn.Module):
    """nn.Linear look-alike whose weight / bias are the first `out_features` rows of the engine's head arena."""

    def __init__(self, eng: ResNetEngine, out_features: int):
        super().__init__()
        self.in_features, self.out_features = eng.feat_dim, out_features
        w, b = eng.fc_views(out_features)
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(b)

    def forward(self, feat):
        return torch.nn.functional.linear(feat, self.weight, self.bias)

    def _apply(self, fn, recurse=True):
        return self


class _Network(nn.Module):
    """`Model` of ewc.py:42-57 / icarl.py:24-38: backbone + classifier."""

    def __init__(self, backbone: CifarResNet, head: _Head):
        super().__init__()
        self.backbone = backbone
        self.classifier = head
        self.feat_dim, self.num_class = head.in_features, head.out_features

    def forward(self, x):
        return self.classifier(self.backbone(x)["features"])

    get_logits = forward


def _draw_linear(in_f: int, out_f: int):
    """Draws an nn.Linear exactly as the reference does (same torch RNG consumption) and returns (weight, bias)."""
    m = nn.Linear(in_f, out_f)
    return m.weight.data, m.bias.data


def _raw_task_samples(dataset):
    """Every item of a task dataset, untransformed, in dataset order: `(images, labels)` tensors.  Tensor-backed datasets expose `.images` / `.labels`
    (the reference's datasets hold image PATHS under the same attribute names, core/data/dataset.py:232-266); anything else is indexed item by item
    with its transform switched off for the duration."""
    imgs, labels = getattr(dataset, "images", None), getattr(dataset, "labels", None)
    if torch.is_tensor(imgs) and labels is not None:
        return imgs, torch.as_tensor(labels, dtype=torch.int64)
    saved = getattr(dataset, "trfms", None)
    if saved is not None:
        dataset.trfms = None
    try:
        items = [dataset[i] for i in range(len(dataset))]
    finally:
        if saved is not None:
            dataset.trfms = saved
    return torch.stack([torch.as_tensor(d["image"]) for d in items]), torch.as_tensor([int(d["label"]) for d in items], dtype=torch.int64)


class _ResNetMethod(nn.Module):
    def __init__(self, backbone, feat_dim, num_class, **kwargs):
        super().__init__()
        if not isinstance(backbone, CifarResNet):
            raise TypeError("libcontinual_b200 methods need a libcontinual_b200 backbone (e.g. cifar_resnet32); "
                            "there is no eager-PyTorch fallback path")
        assert feat_dim == backbone.out_dim
        self.backbone = backbone
        self.engine: ResNetEngine = backbone.engine
        self.feat_dim, self.num_class = feat_dim, num_class
        self.device = kwargs.get("device", self.engine.device)
        self.kwargs = kwargs
        self.task_idx = 0
        assert num_class <= self.engine.cap, "head capacity (backbone kwarg num_classes) too small"

    # -- head management -------------------------------------------------------------------------------------------------
    def _set_head(self, weight: torch.Tensor, bias: torch.Tensor):
        eng = self.engine
        n = weight.shape[0]
        w, b = eng.fc_views(n)
        w.copy_(weight)
        b.copy_(bias)
        eng.ncls = n
        self.network = _Network(self.backbone, _Head(eng, n))

    def _grow_head(self, n_new: int):
        """new Linear whose first rows are the old head (ewc.py:71-80, lwf.py:28-42); fresh rows use the reference's init."""
        eng = self.engine
        n_old = eng.ncls
        w_new, b_new = _draw_linear(self.feat_dim, n_new)
        w, b = eng.fc_views(n_new)
        w[n_old:].copy_(w_new[n_old:])
        b[n_old:].copy_(b_new[n_old:])
        eng.ncls = n_new
        self.network = _Network(self.backbone, _Head(eng, n_new))

    def _params(self):
        # arena-backed parameters only (resnet18 also registers the reference's unused `fc`, resnet.py:191, which never receives a gradient)
        return [p for n, p in self.backbone.named_parameters() if n in self.engine.param_off] + [self.network.classifier.weight, self.network.classifier.bias]

    def _grad_views(self, arena=None):
        eng = self.engine
        arena = eng.grads if arena is None else arena
        gw, gb = eng.fc_views(eng.ncls, arena)
        return tuple(eng.param_view(n, arena) for n, _ in eng.layout) + (gw, gb)

    def get_parameters(self, config):
        return [{"params": self._params()}]

    # -- shared step pieces --------------------------------------------------------------------------------------------------
    def _to_device(self, data):
        x = data["image"].to(self.engine.device, dtype=torch.float32, non_blocking=True).contiguous()
        y = data["label"].to(self.engine.device, dtype=torch.int64, non_blocking=True).contiguous()
        return x, y

    def _params_cached(self):
        """`_params()` walks named_parameters() of the whole backbone (0.4 ms of host time per step); the list only changes when the head does."""
        key = (self.engine.ncls, id(self.network.classifier), getattr(self, "n_old_rows", None))
        c = self.__dict__.get("_params_cache")
        if c is None or c[0] != key:
            c = (key, self._params())
            self.__dict__["_params_cache"] = c
        return c[1]

    def _grad_views_fast(self, arena):
        """`_grad_views(arena)` with the slicing plan cached: one split_with_sizes + one view per tensor instead of a dict lookup, a slice and a view
        each (fresh tensor objects every call, so that autograd can adopt them as p.grad without a copy)."""
        key = (self.engine.ncls, getattr(self, "n_old_rows", None))
        c = self.__dict__.get("_gv_plan")
        if c is None or c[0] != key:
            ref = self._grad_views(arena)
            base = arena.data_ptr()
            offs = [(v.data_ptr() - base) // 4 for v in ref]
            sizes, shapes, pos = [], [], 0
            if any(o2 < o1 + v1.numel() for o1, o2, v1 in zip(offs, offs[1:], ref)):
                return ref                       # not in ascending, disjoint arena order: keep the plain path
            for o, v in zip(offs, ref):          # views are in arena order; gaps (alignment padding, unused head rows) become their own chunks
                if o > pos:
                    sizes.append(o - pos); shapes.append(None)
                sizes.append(v.numel()); shapes.append(tuple(v.shape)); pos = o + v.numel()
            if pos < arena.numel():
                sizes.append(arena.numel() - pos); shapes.append(None)
            c = (key, sizes, shapes)
            self.__dict__["_gv_plan"] = c
        return tuple(ch.view(sh) for ch, sh in zip(arena.split_with_sizes(c[1]), c[2]) if sh is not None)

    def _grad_buffer(self):
        """Persistent flat gradient copy handed to the parameters + its cached per-parameter views (never the live arena: the next fused step overwrites
        that one while p.grad may still be read, clipped in place or stepped on)."""
        eng = self.engine
        key = (eng.ncls, getattr(self, "n_old_rows", None), eng.grads.data_ptr(), eng.grads.numel())
        c = self.__dict__.get("_grad_buf")
        if c is None or c[0] != key:
            buf = torch.empty_like(eng.grads)
            c = (key, buf, self._grad_views(buf))
            self.__dict__["_grad_buf"] = c
        return c[1], c[2]

    def _finish(self, B, y):
        eng = self.engine
        import os
        if os.environ.get("LC_B200_AUTOGRAD_VIEWS") == "1":
            loss = _ArenaLoss.apply(self, eng.scal[0], *self._params_cached())
        else:
            anchor = self.__dict__.get("_anchor")
            if anchor is None:
                anchor = self.__dict__["_anchor"] = torch.zeros(1, device=eng.device, requires_grad=True)
            loss = _ArenaLoss.apply(self, eng.scal[0], anchor)
        pred = eng.pred[:B].clone()
        acc = eng.scal[1].item()                 # the reference syncs here too (finetune.py:24)
        return pred, acc / B, loss

    def _fused_step(self, x, y, *, ce_lo, ce_hi, pred_n, teacher=None, kd_n=0, kd_w=0.0, head_rows=None):
        eng = self.engine
        B = x.shape[0]
        n = eng.ncls
        tl = None
        if teacher is not None and kd_n > 0:
            # the frozen teacher's forward is independent of the student's: it runs on a second stream (forked / joined with events, so the pair
            # is captured as two parallel branches of the step's CUDA graph) — every kernel of a 128-image ResNet32 is far too small to fill 148 SMs
            main = torch.cuda.current_stream(eng.device)
            side = self._teacher_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                tl = eng.teacher_logits(teacher, x)
        eng.forward(x, train=True, update_running=True)
        self.backbone.num_batches_pending += 1      # host bookkeeping of BN.num_batches_tracked (never read by the kernels)
        eng.head_forward(B, n)
        if tl is not None:
            torch.cuda.current_stream(eng.device).wait_stream(side)
        eng.loss(y, B, ce_lo, ce_hi, pred_n, teacher_logits=tl, kd_n=kd_n, kd_w=kd_w, T=2.0)
        eng.head_backward(B, n)
        eng.backward(x)
        self._project_gradients()

    # -- observe() as a CUDA-graph replay ---------------------------------------------------------------------------------------
    def _graph_key(self):
        eng = self.engine
        return (self.task_idx, eng.ncls, getattr(eng, "precision", None), id(getattr(self, "fisher", None)), id(getattr(self, "teacher", None)),
                id(getattr(self, "ref_model", None)), getattr(self, "cur_lamda", None), id(getattr(self, "_gpm", None)))

    def _observe_launch(self, x, y):
        """The kernels of one `observe` (forward, loss, backward, regulariser).  A Trainer calls observe with the same shapes thousands of times per task
        (trainer.py:563-614): the first two calls of a configuration launch eagerly (~200 launches through ctypes), the third captures the same launch
        sequence into a CUDA graph and every later call is a copy into the graph's input buffers plus one replay — the plugin surface stays exactly the
        reference's, the per-launch host work disappears.  LC_B200_EAGER_OBSERVE=1 keeps every call eager."""
        import os
        if os.environ.get("LC_B200_EAGER_OBSERVE") == "1" or not self.training:
            return self._launch_step(x, y)
        obs = self.__dict__.setdefault("_obs_graphs", {})
        key = (x.shape[0],) + self._graph_key()
        st = obs.get(key)
        if st is None:
            if obs and next(iter(obs))[1:] != key[1:]:
                obs.clear()                               # a new task / head size: the old task's graphs (and their memory pool) are released
            st = obs[key] = {"n": 0}
        st["n"] += 1
        if st["n"] <= 2:
            return self._launch_step(x, y)
        if "g" not in st:
            st["x"], st["y"] = torch.empty_like(x), torch.empty_like(y)
            pending = self.backbone.num_batches_pending
            torch.cuda.synchronize(self.engine.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch_step(st["x"], st["y"])
            self.backbone.num_batches_pending = pending   # capture runs the host bookkeeping once without executing a kernel
            st["g"] = g
        st["x"].copy_(x, non_blocking=True)
        st["y"].copy_(y, non_blocking=True)
        st["g"].replay()
        self.backbone.num_batches_pending += 1

    # -- GPM-style gradient projection on top of any ResNet method (BASELINE config C5: "LwF + GPM") ---------------------------------------
    def set_gradient_projection(self, bases):
        """bases: {parameter name (as in `engine.layout`): U [Cin*k*k, r]} — after every backward the named conv gradients lose their component inside
        span(U): g <- g - g.view(Cout, -1) @ (U U^T)  (gpm.py:78-81), in place in the gradient arena through `GPMProjector` (three BF16 tcgen05 GEMMs on
        a two-term split per layer, fp32-level accuracy).  `None` removes it.  The reference has no recipe that combines LwF with GPM (its GPM is
        AlexNet-only); this is the operator C5 names, applied where the reference's GPM applies it: between backward and step."""
        if not bases:
            self._gpm = None
            return
        from ..gpm import GPMProjector
        names = list(bases.keys())
        for n in names:
            shape = self.engine.param_off[n][1]
            assert len(shape) == 4 and (shape[1] * shape[2] * shape[3]) % 8 == 0, f"{n}: Cin*k*k must be a multiple of 8 for the tcgen05 projection"
        self._gpm = (names, GPMProjector([bases[n] for n in names], device=self.engine.device))
        self.__dict__.get("_obs_graphs", {}).clear()

    def _project_gradients(self):
        gp = getattr(self, "_gpm", None)
        if gp is None:
            return
        names, proj = gp
        eng = self.engine
        for i, n in enumerate(names):
            g = eng.param_view(n, eng.grads)
            proj.project_(i, g.view(g.shape[0], -1))
        eng.launches += 4 * len(names)

    def _teacher_stream(self):
        if getattr(self, "_tstream", None) is None:
            self._tstream = torch.cuda.Stream(device=self.engine.device)
        return self._tstream

    def _infer_logits(self, x):
        eng = self.engine
        B = x.shape[0]
        eng.forward(x, train=self.training, update_running=self.training)
        if self.training:
            self.backbone._bump_num_batches_tracked()
        eng.head_forward(B, eng.ncls)
        return eng.logits[:B, :eng.ncls]

    def forward(self, x):
        return self._infer_logits(x.to(self.engine.device, torch.float32).contiguous()).clone()

    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.task_idx = task_idx

    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        pass

    def inference(self, data):
        x, y = self._to_device(data)
        logit = self._infer_logits(x)
        pred = torch.argmax(logit, dim=1)
        return pred, host_acc(self, torch.sum(pred == y), x.size(0))


class Finetune(_ResNetMethod):
    """core/model/finetune.py:4-51."""

    def __init__(self, backbone, feat_dim, num_class, **kwargs):
        super().__init__(backbone, feat_dim, num_class, **kwargs)
        self._set_head(*_draw_linear(feat_dim, num_class))

    def _launch_step(self, x, y):
        n = self.engine.ncls
        self._fused_step(x, y, ce_lo=0, ce_hi=n, pred_n=n)

    def observe(self, data):
        x, y = self._to_device(data)
        self._observe_launch(x, y)
        return self._finish(x.shape[0], y)


class EWC(_ResNetMethod):
    """core/model/ewc.py:59-229.  State: theta* and Fisher as flat arenas with the parameter-arena layout."""

    def __init__(self, backbone, feat_dim, num_class, **kwargs):
        super().__init__(backbone, feat_dim, num_class, **kwargs)
        _draw_linear(feat_dim, num_class)                          # Finetune.classifier (finetune.py:10): unused, but draws RNG
        self._set_head(*_draw_linear(feat_dim, kwargs["init_cls_num"]))
        self.lamda = kwargs["lamda"]
        self.ref_param = self.engine.params.clone()
        self.fisher = torch.zeros_like(self.engine.params)
        self.fisher_rows = kwargs["init_cls_num"]                  # len(self.fisher['classifier.weight'])

    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.task_idx = task_idx
        self._grow_head(self.kwargs["init_cls_num"] + task_idx * self.kwargs["inc_cls_num"])

    def _launch_step(self, x, y):
        eng = self.engine
        n = eng.ncls
        if self.task_idx == 0:
            self._fused_step(x, y, ce_lo=0, ce_hi=n, pred_n=n)
        else:
            old = n - self.kwargs["inc_cls_num"]
            self._fused_step(x, y, ce_lo=old, ce_hi=n, pred_n=n)
            eng.ewc_penalty(self.ref_param, self.fisher, float(self.lamda))

    def observe(self, data):
        x, y = self._to_device(data)
        self._observe_launch(x, y)
        return self._finish(x.shape[0], y)

    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        """theta* <- theta; Fisher pass over the task loader in train() mode, CE over ALL logits (ewc.py:110-133,147-205)."""
        eng = self.engine
        self.ref_param = eng.params.clone()
        new_fisher = torch.zeros_like(eng.params)
        n = eng.ncls
        nb = 0
        for data in train_loader:
            x, y = self._to_device(data)
            self._fused_step(x, y, ce_lo=0, ce_hi=n, pred_n=n)
            eng.fisher_accumulate(new_fisher, float(x.shape[0]))
            nb += 1
        num_samples = float(train_loader.batch_size * len(train_loader))       # ewc.py:202 (over-counts a ragged last batch)
        alpha = 1 - self.kwargs["inc_cls_num"] / n
        lib, st = eng.lib, torch.cuda.current_stream().cuda_stream
        from .._lib import check
        r, fd = self.fisher_rows, eng.feat_dim
        segs = [(0, eng.n_backbone, True), (eng.off_fc_w, r * fd, True), (eng.off_fc_w + r * fd, (n - r) * fd, False),
                (eng.off_fc_b, r, True), (eng.off_fc_b + r, n - r, False)]
        for off, cnt, merge in segs:
            if cnt > 0:
                check(lib.lc_fisher_merge(new_fisher.data_ptr() + 4 * off, (self.fisher.data_ptr() + 4 * off) if merge else None, cnt, num_samples,
                                          float(alpha), st), "lc_fisher_merge")
        self.fisher = new_fisher
        self.fisher_rows = n


class ICarl(_ResNetMethod):
    """core/model/icarl.py:42-221 (training / distillation path; exemplar management is host-side)."""

    def __init__(self, backbone, feat_dim, num_class, **kwargs):
        super().__init__(backbone, feat_dim, num_class, **kwargs)
        self._set_head(*_draw_linear(feat_dim, num_class))
        self.cur_task_id = 0
        self.cur_cls_indexes = None
        self.old_network = None
        self.prev_cls_num = 0
        self.accu_cls_num = 0
        self.init_cls_num, self.inc_cls_num, self.task_num = kwargs["init_cls_num"], kwargs["inc_cls_num"], kwargs["task_num"]
        self.class_means = None

    def get_parameters(self, config):
        return self._params()

    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.accu_cls_num = self.init_cls_num if self.cur_task_id == 0 else self.accu_cls_num + self.inc_cls_num
        self.cur_cls_indexes = np.arange(self.prev_cls_num, self.accu_cls_num)

    def snapshot_teacher(self):
        """The teacher lines of after_task (icarl.py:172-176): frozen eval-mode copy of the current network."""
        self.old_network = TeacherState(self.engine)
        self.prev_cls_num = self.accu_cls_num

    def after_task(self, task_idx, buffer, train_loader, test_loaders):
        """icarl.py:169-191: freeze the teacher, shrink + refill the exemplar memory by herding, recompute the class means."""
        self.snapshot_teacher()
        val_transform = None
        if test_loaders:
            val_transform = getattr(test_loaders[0].dataset, "trfms", None)
        if buffer is not None and hasattr(buffer, "herding_indices") and train_loader is not None:
            # tensor-backed memory (libcontinual_b200.buffer.HerdingBuffer).  Like linearherdingbuffer.py:94-111 the pool is the task's DATASET, not
            # the training loader: raw items in dataset order (no shuffle, nothing dropped, no training augmentation), replayed exemplars removed,
            # features taken under the validation transform; the raw items are what gets stored.
            x, y = _raw_task_samples(train_loader.dataset)
            keep = (y >= int(self.cur_cls_indexes[0])) & (y <= int(self.cur_cls_indexes[-1]))
            x, y = x[keep], y[keep]
            order = torch.sort(y, stable=True)[1]             # remove_buffer_sample_in_dataset regroups by class, order kept inside a class
            buffer.reduce_old_data(self.cur_task_id, self.accu_cls_num)
            buffer.update(self, x[order], y[order], self.accu_cls_num, transform=val_transform)
            self.class_means = buffer.class_means(self, transform=val_transform).to(self.engine.device).contiguous()
        elif buffer is not None and hasattr(buffer, "reduce_old_data"):
            # a reference-style path buffer (core/model/buffer/linearherdingbuffer.py): its own update, then icarl.py:187-190
            buffer.reduce_old_data(self.cur_task_id, self.accu_cls_num)
            buffer.update(self.network, train_loader, val_transform, self.cur_task_id, self.accu_cls_num, self.cur_cls_indexes, self.device)
            self.class_means = self.calc_class_mean(buffer, train_loader, val_transform, self.device).to(self.engine.device).contiguous()
        self.cur_task_id += 1

    @torch.no_grad()
    def calc_class_mean(self, buffer, train_loader, val_transform, device):
        """icarl.py:226-287 for a buffer of image paths: decode every stored exemplar under the validation transform, per class the mean of the
        L2-normalised backbone features, re-normalised."""
        import os
        import PIL.Image
        ds = train_loader.dataset
        root, mode = ds.data_root, ds.mode
        was = self.backbone.training
        self.backbone.eval()
        feats, labels = [], []
        bs = train_loader.batch_size or 32
        for i in range(0, len(buffer.labels), bs):
            imgs = [val_transform(PIL.Image.open(os.path.join(root, mode, pth)).convert("RGB")) for pth in buffer.images[i:i + bs]]
            f = self.backbone(torch.stack(imgs))["features"]
            feats.append((f / f.norm(dim=1).view(-1, 1)).cpu())
            labels.extend(int(l) for l in buffer.labels[i:i + bs])
        self.backbone.train(was)
        feats, labels = torch.cat(feats), np.asarray(labels)
        means = []
        for c in np.unique(labels):
            m = feats[np.where(labels == c)[0]].mean(0)
            means.append(m / m.norm())
        return torch.stack(means)

    def _launch_step(self, x, y):
        kd = self.cur_task_id > 0 and self.old_network is not None
        self._fused_step(x, y, ce_lo=0, ce_hi=self.accu_cls_num, pred_n=self.accu_cls_num, teacher=self.old_network if kd else None,
                         kd_n=self.prev_cls_num if kd else 0, kd_w=1.0)

    def observe(self, data):
        x, y = self._to_device(data)
        self._observe_launch(x, y)
        return self._finish(x.shape[0], y)

    def inference(self, data):
        """icarl.py:96-152: nearest-class-mean once the class means of every seen class exist, else the linear head."""
        x, y = self._to_device(data)
        eng = self.engine
        if self.class_means is not None and len(self.class_means) == self.accu_cls_num:
            B = x.shape[0]
            eng.forward(x, train=self.training, update_running=self.training)
            eng.pool_forward(B)
            from .._lib import check
            check(eng.lib.lc_ncm_classify(eng.features(B).data_ptr(), self.class_means.data_ptr(), B, self.accu_cls_num, eng.feat_dim, eng.pred.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream), "lc_ncm_classify")
            pred = eng.pred[:B].clone()
        else:
            logits = self._infer_logits(x)[:, :self.accu_cls_num]
            pred = torch.argmax(logits, dim=1)
        return pred, host_acc(self, torch.sum(pred == y), x.size(0))


class LWF(_ResNetMethod):
    """core/model/lwf.py:9-81 (lamda = 3 and T = 2 are hard-coded by the reference at :63-65)."""

    def __init__(self, backbone, feat_dim, num_class, **kwargs):
        super().__init__(backbone, feat_dim, num_class, **kwargs)
        _draw_linear(feat_dim, num_class)                          # Finetune.classifier, replaced right away (lwf.py:14)
        self._set_head(*_draw_linear(feat_dim, kwargs["init_cls_num"]))
        self.init_cls_num, self.inc_cls_num = kwargs["init_cls_num"], kwargs["inc_cls_num"]
        self.known_cls_num = 0
        self.total_cls_num = 0
        self.old = None

    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.task_idx = task_idx
        self.known_cls_num = self.total_cls_num
        self.total_cls_num = self.init_cls_num + task_idx * self.inc_cls_num
        eng = self.engine
        # update_fc (lwf.py:28-42): fresh Linear(total), old rows copied over
        n_old = eng.ncls
        w_new, b_new = _draw_linear(self.feat_dim, self.total_cls_num)
        if task_idx != 0:
            self.old = TeacherState(eng)                           # old_fc + old_backbone, frozen, eval (lwf.py:31,49-50)
        w, b = eng.fc_views(self.total_cls_num)
        w[n_old:].copy_(w_new[n_old:])
        b[n_old:].copy_(b_new[n_old:])
        eng.ncls = self.total_cls_num
        self.network = _Network(self.backbone, _Head(eng, self.total_cls_num))

    def _launch_step(self, x, y):
        n = self.engine.ncls
        if self.task_idx == 0:
            self._fused_step(x, y, ce_lo=0, ce_hi=n, pred_n=n)
        else:
            self._fused_step(x, y, ce_lo=self.known_cls_num, ce_hi=n, pred_n=n, teacher=self.old, kd_n=self.known_cls_num, kd_w=3.0)

    def observe(self, data):
        x, y = self._to_device(data)
        self._observe_launch(x, y)
        return self._finish(x.shape[0], y)


class _CosineHead(nn.Module):
    """`CosineLinear` / `SplitCosineLinear` look-alike (resnet.py:418-463) over the engine's head arena: weight rows [0, n_old) are
    `fc1.weight`, rows [n_old, n) `fc2.weight`, `sigma` lives in the first slot of the bias region."""

    class _Part(nn.Module):
        def __init__(self, w):
            super().__init__()
            self.weight = nn.Parameter(w)
            self.out_features, self.in_features = w.shape

        def _apply(self, fn, recurse=True):
            return self

    def __init__(self, eng: ResNetEngine, n_old: int, n_new: int):
        super().__init__()
        n = n_old + n_new
        self.in_features, self.out_features = eng.feat_dim, n
        w, b = eng.fc_views(n)
        self.sigma = nn.Parameter(eng.params[eng.off_fc_b:eng.off_fc_b + 1])
        if n_old == 0:                     # task 0: a plain CosineLinear (one weight); later tasks: SplitCosineLinear (fc1 = old rows, fc2 = new rows)
            self.weight = nn.Parameter(w)
        else:
            self.fc1 = _CosineHead._Part(w[:n_old])
            self.fc2 = _CosineHead._Part(w[n_old:])

    def _apply(self, fn, recurse=True):
        return self


class LUCIR(_ResNetMethod):
    """core/model/lucir.py:72-240 on the `resnet32_V2` backbone: cosine head, less-forget + CE + margin-ranking losses, frozen
    reference model, old-class embedding excluded from the update (param group with lr 0)."""

    def __init__(self, backbone, feat_dim, num_class, **kwargs):
        super().__init__(backbone, feat_dim, num_class, **kwargs)
        _draw_linear(feat_dim, num_class)                           # Finetune.classifier (finetune.py:10): unused, draws RNG
        stdv = 1.0 / (feat_dim ** 0.5)                              # CosineLinear.reset_parameters (resnet.py:431-435)
        w0 = torch.empty(kwargs["init_cls_num"], feat_dim).uniform_(-stdv, stdv)
        eng = self.engine
        self._set_cos_head(w0, 1.0, 0)
        self.K, self.lw_mr, self.dist, self.lamda = kwargs["K"], kwargs["lw_mr"], kwargs["dist"], kwargs["lamda"]
        self.ref_model = None
        self.cur_lamda = self.lamda
        self.num_old_classes = 0

    def _set_cos_head(self, weight, sigma, n_old):
        eng = self.engine
        n = weight.shape[0]
        w, _ = eng.fc_views(n)
        w.copy_(weight)
        eng.params[eng.off_fc_b] = float(sigma)
        eng.ncls = n
        self.n_old_rows = n_old
        self.network = _Network.__new__(_Network)
        nn.Module.__init__(self.network)
        self.network.backbone = self.backbone
        self.network.classifier = _CosineHead(eng, n_old, n - n_old)
        self.network.feat_dim, self.network.num_class = self.feat_dim, n

    def _class_mean_embeddings(self, train_loader, first_cls, n_new):
        """`_init_new_fc` (lucir.py:130-158): eval-mode features of every sample of each new class, L2-normalised, averaged and
        re-normalised in float64 on the host exactly as the reference does with its numpy feature matrix."""
        eng = self.engine
        was_training = self.backbone.training
        self.backbone.eval()
        sums = torch.zeros(n_new, self.feat_dim, dtype=torch.float64)
        counts = torch.zeros(n_new, dtype=torch.float64)
        with torch.no_grad():
            for data in train_loader:
                x, y = self._to_device(data)
                f = self.backbone.feature(x).double().cpu()
                f = torch.nn.functional.normalize(f, p=2, dim=1)
                yl = y.cpu()
                for c in range(n_new):
                    m = yl == first_cls + c
                    if m.any():
                        sums[c] += f[m].sum(0); counts[c] += float(m.sum())
        self.backbone.train(was_training)
        mean = sums / counts.clamp_min(1.0).unsqueeze(1)
        return torch.nn.functional.normalize(mean, p=2, dim=1)

    def before_task(self, task_idx, buffer, train_loader, test_loaders):
        self.task_idx = task_idx
        eng = self.engine
        inc = self.kwargs["inc_cls_num"]
        if task_idx >= 1:
            n_prev = eng.ncls
            self.ref_model = TeacherState(eng)                     # copy.deepcopy(self.network), eval (lucir.py:88,99,121)
            self.num_old_classes = n_prev
            stdv = 1.0 / (self.feat_dim ** 0.5)
            # SplitCosineLinear(in, n_prev, inc): both parts draw their uniform init, fc1 is then overwritten by the old embedding
            torch.empty(n_prev, self.feat_dim).uniform_(-stdv, stdv)
            w2 = torch.empty(inc, self.feat_dim).uniform_(-stdv, stdv)
            old_w, _ = eng.fc_views(n_prev)
            old_w = old_w.clone()
            if train_loader is not None:                           # _init_new_fc: class-mean embedding scaled to the old average norm
                avg_norm = old_w.norm(dim=1, keepdim=True).mean(dim=0).double().cpu()
                emb = self._class_mean_embeddings(train_loader, n_prev, inc)
                w2 = (emb * avg_norm).float()
            sigma = float(eng.params[eng.off_fc_b])
            self._set_cos_head(torch.cat([old_w, w2.to(old_w.device)]), sigma, n_prev)
            self.cur_lamda = self.lamda * ((n_prev * 1.0 / inc) ** 0.5)
        else:
            self.cur_lamda = self.lamda

    def _params(self):
        c = self.network.classifier
        head = [c.weight] if hasattr(c, "weight") else [c.fc1.weight, c.fc2.weight]
        return [p for _, p in self.backbone.named_parameters()] + head + [c.sigma]

    def _grad_views(self, arena=None):
        eng = self.engine
        arena = eng.grads if arena is None else arena
        gw, _ = eng.fc_views(eng.ncls, arena)
        heads = (gw,) if self.n_old_rows == 0 else (gw[:self.n_old_rows], gw[self.n_old_rows:])
        return tuple(eng.param_view(n, arena) for n, _ in eng.layout) + heads + (arena[eng.off_fc_b:eng.off_fc_b + 1],)

    def get_parameters(self, config):
        if self.task_idx > 0:       # lucir.py:229-240 (the lr 0.1 / wd 5e-4 literals are the reference's)
            c = self.network.classifier
            base = [p for _, p in self.backbone.named_parameters()] + [c.fc2.weight, c.sigma]
            return [{"params": base, "lr": 0.1, "weight_decay": 5e-4}, {"params": [c.fc1.weight], "lr": 0, "weight_decay": 0}]
        return self._params()

    def _launch_step(self, x, y):
        eng = self.engine
        lib, st = eng.lib, torch.cuda.current_stream().cuda_stream
        from .._lib import check
        B, n = x.shape[0], eng.ncls
        kd = self.task_idx > 0 and self.ref_model is not None
        ref_feat = eng.features(B)
        if kd:
            t = self.ref_model
            eng.forward(x, train=False, update_running=False, params=t.params, rstat=t.rstat, ws=t.ws)
            eng.pool_forward(B, ws=t.ws)
            ref_feat = eng.features(B, ws=t.ws)
        eng.forward(x, train=True, update_running=True)
        self.backbone.num_batches_pending += 1
        eng.pool_forward(B)
        feat = eng.features(B)
        P = eng.params.data_ptr()
        w_ptr, sig_ptr = P + 4 * eng.off_fc_w, P + 4 * eng.off_fc_b
        check(lib.lc_cosine_head_forward(feat.data_ptr(), w_ptr, sig_ptr, B, n, eng.feat_dim, eng.cos_inv_norm.data_ptr(), eng.scores.data_ptr(),
                                         eng.logits.data_ptr(), eng.cap, st), "lc_cosine_head_forward")
        G = eng.grads.data_ptr()
        check(lib.lc_lucir_loss(eng.logits.data_ptr(), eng.scores.data_ptr(), eng.cap, feat.data_ptr(), ref_feat.data_ptr(), eng.feat_dim, y.data_ptr(), B, n,
                                self.num_old_classes if kd else 0, int(self.K), float(self.cur_lamda if kd else 0.0), float(self.dist),
                                float(self.lw_mr if kd else 0.0), eng.dlogits.data_ptr(), eng.dscores.data_ptr(), eng.dfeat_extra.data_ptr(), eng.pred.data_ptr(),
                                eng.scal.data_ptr(), G + 4 * eng.off_fc_b, st), "lc_lucir_loss")
        # gscores = sigma * dlogits + dscores  (one fused elementwise over [B, cap])
        gs = torch.addcmul(eng.dscores, eng.dlogits, eng.params[eng.off_fc_b:eng.off_fc_b + 1])
        dfeat = eng.ws[eng._off[5]:eng._off[5] + B * eng.feat_dim]
        check(lib.lc_cosine_head_backward(gs.data_ptr(), eng.cap, feat.data_ptr(), w_ptr, eng.cos_inv_norm.data_ptr(), B, n, eng.feat_dim, dfeat.data_ptr(),
                                          G + 4 * eng.off_fc_w, st), "lc_cosine_head_backward")
        dfeat.view(B, eng.feat_dim).add_(eng.dfeat_extra[:B])
        eng.pool_backward(dfeat.view(B, eng.feat_dim))
        eng.backward(x)
        eng.launches += 6

    def observe(self, data):
        x, y = self._to_device(data)
        self._observe_launch(x, y)
        return self._finish(x.shape[0], y)

    def _infer_logits(self, x):
        eng = self.engine
        B = x.shape[0]
        eng.forward(x, train=self.training, update_running=self.training)
        eng.pool_forward(B)
        P = eng.params.data_ptr()
        from .._lib import check
        check(eng.lib.lc_cosine_head_forward(eng.features(B).data_ptr(), P + 4 * eng.off_fc_w, P + 4 * eng.off_fc_b, B, eng.ncls, eng.feat_dim,
                                             eng.cos_inv_norm.data_ptr(), eng.scores.data_ptr(), eng.logits.data_ptr(), eng.cap,
                                             torch.cuda.current_stream().cuda_stream), "lc_cosine_head_forward")
        return eng.logits[:B, :eng.ncls]
