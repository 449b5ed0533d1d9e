"""Step loop of the reference `Trainer._train` (core/trainer.py:563-614) on top of the fused kernels.

Two ways to run one training step of a method plugin:

* `train_step_eager(model, optimizer, batch)` — literally the reference order (`observe` -> `zero_grad` -> `loss.backward()` ->
  `optimizer.step()` -> `loss.item()`), usable with torch.optim or `libcontinual_b200.optim`.
* `GraphedStep` — the same kernel sequence (teacher forward, backbone forward, head, loss, backward, regulariser, fused SGD)
  captured once into a CUDA graph and replayed per batch: no per-launch host work, no autograd bookkeeping, metrics stay on the
  device until asked for.  With `world_size > 1` the flat gradient arena is all-reduced (NCCL, average) between backward and
  the optimizer kernel — one bucket, the only collective on the data path.
"""
from __future__ import annotations

from typing import Optional

import torch

from .optim import SGD, Adam
from .parallel import allreduce_mean_


def _capture_with_collective(launch_fwd_bwd, flat_grads, launch_update, pg, world):
    """One step as CUDA graph(s).  world == 1: one graph.  world > 1: forward/backward, the gradient all-reduce (NCCL average over NVLink/NVSwitch — NCCL
    collectives are capturable) and the optimizer kernels in ONE graph, so a step is a single graph launch with no host work between backward and update
    (at 16 images per GPU the whole step is ~1 ms and a second graph launch plus an eager collective would be a tenth of it).
    `LC_DP_COLLECTIVE=eager` keeps the collective outside (two graphs).  Returns (g_main, g_upd, description)."""
    import os
    g_main = torch.cuda.CUDAGraph()
    if world == 1:
        with torch.cuda.graph(g_main):
            launch_fwd_bwd()
            launch_update()
        return g_main, None, "none (single replica)"
    if os.environ.get("LC_DP_COLLECTIVE", "graph") != "eager":
        allreduce_mean_(flat_grads, pg)                       # communicator and NCCL buffers exist before capture starts
        torch.cuda.synchronize()
        with torch.cuda.graph(g_main):
            launch_fwd_bwd()
            allreduce_mean_(flat_grads, pg)
            launch_update()
        return g_main, None, f"ncclAvg all-reduce of the flat gradient arena ({flat_grads.numel() * 4} B, one bucket) captured inside the step's CUDA graph"
    with torch.cuda.graph(g_main):
        launch_fwd_bwd()
    g_upd = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g_upd):
        launch_update()
    return g_main, g_upd, f"eager ncclAvg all-reduce of the flat gradient arena ({flat_grads.numel() * 4} B, one bucket) between two CUDA graphs"


def train_step_eager(model, optimizer, batch):
    """trainer.py:601-612, default branch."""
    pred, acc, loss = model.observe(batch)
    optimizer.zero_grad()
    loss.backward()
    optimizer.step()
    return pred, acc, loss.item()


class GraphedStep:
    def __init__(self, model, optimizer: SGD, batch_size: int, process_group=None, warmup: int = 3):
        eng = model.engine
        assert isinstance(optimizer, SGD), "GraphedStep drives the fused flat optimizer"
        self.model, self.opt, self.eng, self.B = model, optimizer, eng, batch_size
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (process_group is not None or (
            torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        dev = eng.device
        self.x = torch.zeros(batch_size, eng.in_ch, eng.img, eng.img, device=dev)
        self.y = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self._sync_hp()
        # param groups beyond the first must be frozen ranges (LUCIR's old-class embedding, lucir.py:229-240): the captured update skips them exactly
        # like the eager `SGD.step` does; anything else raises here instead of silently training with group 0's hyper-parameters
        self._frozen = optimizer._frozen_range()
        # warm-up on a side stream (sets func attributes, primes the allocator), then capture
        pending0 = model.backbone.num_batches_pending
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        saved = (eng.params.clone(), eng.rstat.clone(), optimizer.buf.clone())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._fwd_bwd()
                self._update()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        eng.params.copy_(saved[0]); eng.rstat.copy_(saved[1]); optimizer.buf.copy_(saved[2])
        l0 = eng.launches
        self.g_main, self.g_upd, self.collective = _capture_with_collective(self._fwd_bwd, eng.grads, self._update, self.pg, self.world)
        self.launches_per_step = eng.launches - l0 + (1 if self.world > 1 else 0)
        model.backbone.num_batches_pending = pending0      # warm-up / capture launches restored above do not count
        self.steps = 0
        self.stager = _InputStager(self.x, self.y)

    def prefetch(self, x: torch.Tensor, y: torch.Tensor):
        """Optional: start the host -> device copy of the batch that the NEXT `run` will be given (copy stream, double-buffered): the 1.6 MB of a
        128-image CIFAR batch is 60 us of PCIe time, 4 % of a step, that otherwise sits in front of every replay."""
        self.stager.prefetch(x, y)

    def _sync_hp(self):
        g = self.opt.param_groups[0]
        hp = (float(g["lr"]), float(g["momentum"]), float(g["weight_decay"]))
        if hp != self.opt._hp_host:
            self.opt.hp[:3] = torch.tensor(hp, device=self.opt.hp.device)
            self.opt._hp_host = hp

    def _fwd_bwd(self):
        self.model._launch_step(self.x, self.y)

    def _update(self):
        self.eng.sgd_step(self.opt.buf, self.opt.hp, frozen=self._frozen)

    def run(self, x: torch.Tensor, y: torch.Tensor, non_blocking: bool = True):
        """One step on a batch that is either device-resident or in (pinned) host memory.  Returns nothing: read
        `loss()` / `correct()` when needed (device scalars)."""
        self._sync_hp()
        self.stager.load(x, y, self.x, self.y, non_blocking)
        self.g_main.replay()
        if self.g_upd is not None:
            allreduce_mean_(self.eng.grads, self.pg)
            self.g_upd.replay()
        self.steps += 1
        self.model.backbone.num_batches_pending += 1

    def loss(self) -> torch.Tensor:
        return self.eng.scal[0]

    def correct(self) -> torch.Tensor:
        return self.eng.scal[1]


class _InputStager:
    """Double-buffered host -> device staging on a copy stream: `prefetch(x, y)` starts the H2D copy of the NEXT batch while the current step computes;
    `run(x, y)` of the same tensors then only does a device-to-device copy into the graph's input buffers (77 MB of images per ViT step is 1.5-3 ms of
    PCIe time that would otherwise sit in front of every step)."""

    def __init__(self, x: torch.Tensor, y: torch.Tensor):
        dev = x.device
        self.copy = torch.cuda.Stream(device=dev)
        self.sx = [torch.empty_like(x), torch.empty_like(x)]
        self.sy = [torch.empty_like(y), torch.empty_like(y)]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.free = [torch.cuda.Event(), torch.cuda.Event()]
        self.k = 0
        self.pending = None

    def prefetch(self, x: torch.Tensor, y: torch.Tensor):
        k = self.k ^ 1
        self.copy.wait_event(self.free[k])              # the device-to-device copy that last read staging buffer k
        with torch.cuda.stream(self.copy):
            self.sx[k].copy_(x, non_blocking=True)
            self.sy[k].copy_(y, non_blocking=True)
            self.ready[k].record(self.copy)
        self.pending = (k, x.data_ptr(), y.data_ptr())

    def load(self, x: torch.Tensor, y: torch.Tensor, dst_x: torch.Tensor, dst_y: torch.Tensor, non_blocking: bool = True):
        if self.pending is not None and self.pending[1] == x.data_ptr() and self.pending[2] == y.data_ptr():
            k = self.pending[0]
            main = torch.cuda.current_stream(dst_x.device)
            main.wait_event(self.ready[k])
            dst_x.copy_(self.sx[k]); dst_y.copy_(self.sy[k])
            self.free[k].record(main)
            self.k, self.pending = k, None
        else:
            dst_x.copy_(x, non_blocking=non_blocking)
            dst_y.copy_(y, non_blocking=non_blocking)


class GraphedL2PStep:
    """The L2P step (query pass, prompt selection, prompted pass, masked loss, backward to the prompt rows, clip, Adam) as CUDA graphs.
    world_size > 1: graph 1 ends before the clip, the flat trainable-gradient arena (123k floats) is all-reduced (average), graph 2
    clips and updates — the reference's DDP order (gradients averaged in backward, then l2p.py:104, then optimizer.step).
    Also drives the other flat-arena methods trained with Adam (DualPrompt, CodaPrompt): they have no clip stage (`_launch_clip` absent)."""

    def __init__(self, model, optimizer: Adam, batch_size: int, process_group=None, warmup: int = 2):
        assert isinstance(optimizer, Adam), "GraphedL2PStep drives the fused flat Adam"
        self.model, self.opt, self.eng, self.B = model, optimizer, model.engine, batch_size
        self.has_clip = hasattr(model, "_launch_clip")
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (process_group is not None or (
            torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        dev = self.eng.dev
        self.x = torch.zeros(batch_size, 3, 224, 224, device=dev)
        self.y = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        self._hp_static = None
        self._t_dev = None
        self._stage_hp()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        saved = (model.theta.clone(), optimizer.m.clone(), optimizer.v.clone())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                model._launch_step(self.x, self.y, clip=True)
                optimizer.launch(tick=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.theta.copy_(saved[0]); optimizer.m.copy_(saved[1]); optimizer.v.copy_(saved[2])
        self._t_dev = None
        self._stage_hp()
        l0 = self.eng.launches

        def upd():
            if self.has_clip:
                model._launch_clip()               # non-linear in g: runs on the REDUCED gradient (l2p.py:104 after DDP's averaging)
            optimizer.launch(tick=True)
        self.g_main, self.g_upd, self.collective = _capture_with_collective(lambda: model._launch_step(self.x, self.y, clip=False), model.theta_grad, upd,
                                                                            self.pg, self.world)
        self.launches_per_step = self.eng.launches - l0 + 2 + (1 if self.world > 1 else 0)
        self.steps = 0
        self.stager = _InputStager(self.x, self.y)

    def prefetch(self, x: torch.Tensor, y: torch.Tensor):
        """Optional: start the host -> device copy of the batch that the NEXT `run` will be given."""
        self.stager.prefetch(x, y)

    def _stage_hp(self):
        """Hyper-parameters reach the device only when they CHANGE (scheduler step), through a pageable-source copy that is complete for the host
        when it returns; the step count and both bias corrections live on the device (`lc_adam_tick` inside the graph), so a replay that runs
        behind the host can never see a later step's values (the race a re-used pinned staging buffer would have)."""
        h = self.opt.hyper(max(self.opt.t, 1))
        if self._hp_static != h[:5]:
            self.opt.hp[:5].copy_(torch.tensor(h[:5], dtype=torch.float32))
            self.opt.hp[8:12].view(torch.float64).copy_(torch.tensor(h[1:3], dtype=torch.float64))     # betas as doubles for the device-side 1 - beta^t
            self._hp_static = h[:5]
        if self._t_dev != self.opt.t:                    # (re)seed the device counter: construction, or eager steps taken in between
            self.opt.hp[7:8].copy_(torch.tensor([float(self.opt.t)], dtype=torch.float32))
            self._t_dev = self.opt.t

    def run(self, x: torch.Tensor, y: torch.Tensor, non_blocking: bool = True):
        self._stage_hp()
        self.opt.t += 1
        self._t_dev = self.opt.t                         # the graph's tick kernel advances the device counter
        self.stager.load(x, y, self.x, self.y, non_blocking)
        self.g_main.replay()
        if self.g_upd is not None:
            allreduce_mean_(self.model.theta_grad, self.pg)
            self.g_upd.replay()
        self.steps += 1

    def loss(self) -> torch.Tensor:
        return self.model.scal[0]

    def correct(self) -> torch.Tensor:
        return self.model.scal[1]


class GraphedFlatStep:
    """A flat-arena method (`model.theta` / `model.theta_grad` / `model._launch_step(x, y)`, e.g. `model.inflora.InfLoRA_OPT`) with
    `optim.FlatSGD`: forward + backward captured as one CUDA graph, the optimizer kernels as a second one; with world_size > 1 the flat
    gradient arena is all-reduced (average) in between — the only collective on the data path."""

    def __init__(self, model, optimizer, batch_size: int, process_group=None, warmup: int = 2, img: int = 224):
        from .optim import FlatSGD
        assert isinstance(optimizer, FlatSGD), "GraphedFlatStep drives the fused flat SGD"
        self.model, self.opt, self.eng, self.B = model, optimizer, model.engine, batch_size
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if (process_group is not None or (
            torch.distributed.is_available() and torch.distributed.is_initialized())) else 1
        dev = self.eng.dev
        self.x = torch.zeros(batch_size, 3, img, img, device=dev)
        self.y = torch.zeros(batch_size, dtype=torch.int64, device=dev)
        optimizer.sync_hp()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        saved = (model.theta.clone(), optimizer.buf.clone())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                model._launch_step(self.x, self.y)
                optimizer.launch()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        model.theta.copy_(saved[0]); optimizer.buf.copy_(saved[1])
        l0 = self.eng.launches
        self.g_main, self.g_upd, self.collective = _capture_with_collective(lambda: model._launch_step(self.x, self.y), model.theta_grad, optimizer.launch,
                                                                            self.pg, self.world)
        self.launches_per_step = self.eng.launches - l0 + len(model.active_ranges()) + (1 if self.world > 1 else 0)
        self.steps = 0
        self.stager = _InputStager(self.x, self.y)

    def prefetch(self, x: torch.Tensor, y: torch.Tensor):
        """Optional: start the host -> device copy of the batch that the NEXT `run` will be given."""
        self.stager.prefetch(x, y)

    def run(self, x: torch.Tensor, y: torch.Tensor, non_blocking: bool = True):
        self.opt.sync_hp()
        self.stager.load(x, y, self.x, self.y, non_blocking)
        self.g_main.replay()
        if self.g_upd is not None:
            allreduce_mean_(self.model.theta_grad, self.pg)
            self.g_upd.replay()
        self.steps += 1

    def loss(self) -> torch.Tensor:
        return self.model.scal[0]

    def correct(self) -> torch.Tensor:
        return self.model.scal[1]


# ---------------------------------------------------------------------------------------------------------------------------------------------------
# Evaluation loop: `Trainer._validate` (core/trainer.py:616-720)
# ---------------------------------------------------------------------------------------------------------------------------------------------------
class EvalMeter:
    """#correct / #seen per task, accumulated on the device by `lc_eval_meter` (integer atomics) and read back ONCE per validation pass.  The reference
    converts every batch's accuracy to a Python float (`.item()` inside `inference`, then `int(acc * batch_size)`, trainer.py:644-645) — a device
    synchronisation per batch that bounds small-model evaluation once the step itself is a graph replay."""

    def __init__(self, ntask: int, device, bounds=None):
        from . import _lib
        self.lib = _lib.load()
        self.ntask = int(ntask)
        self.counts = torch.zeros(self.ntask, 2, dtype=torch.int64, device=device)       # uint64 on the device; values stay far below 2^63
        self.batch = torch.zeros(self.ntask, 2, dtype=torch.int64, device=device)        # the current batch's counts, folded into `counts` per batch
        self.bounds = None if bounds is None else torch.tensor(list(bounds), dtype=torch.int32, device=device)
        assert self.bounds is None or self.bounds.numel() == self.ntask + 1

    def update(self, pred: torch.Tensor, label: torch.Tensor, task: int = -1, reference_rounding: bool = False):
        """reference_rounding: fold the batch as `int(acc * batch_size)` (trainer.py:644, per-task mode) instead of the exact count."""
        from ._lib import check
        assert pred.is_cuda and label.is_cuda and pred.dtype == torch.int64 and label.dtype == torch.int64 and pred.is_contiguous() and label.is_contiguous()
        st = torch.cuda.current_stream(pred.device).cuda_stream
        check(self.lib.lc_eval_meter(pred.data_ptr(), label.data_ptr(), int(pred.numel()), None if self.bounds is None else self.bounds.data_ptr(), self.ntask,
                                     int(task), self.batch.data_ptr(), st), "lc_eval_meter")
        check(self.lib.lc_eval_fold(self.batch.data_ptr(), self.counts.data_ptr(), self.ntask, int(reference_rounding), st), "lc_eval_fold")

    def result(self):
        c = self.counts.cpu().numpy()                     # the one synchronisation of the pass
        return c[:, 0].copy(), c[:, 1].copy()


@torch.no_grad()
def validate(model, dataloaders, task_idx: int, *, setting: str = "task-agnostic", testing_per_task: bool = True, init_cls_num: Optional[int] = None,
             inc_cls_num: Optional[int] = None):
    """`Trainer._validate(task_idx)` (trainer.py:616-720) -> {'avg_acc', 'per_task_acc'} with the same rounding (`round(100 * correct / count, 2)`).

    dataloaders: the per-task test loaders of tasks 0..task_idx (iterables of {'image', 'label'} batches; `libcontinual_b200.data.GpuLoader` keeps them on
    the device).  testing_per_task=True runs them one after another (task-aware: `inference(batch, task_id=t)`); False evaluates all batches and attributes
    every sample to the task whose class range holds its label (needs init_cls_num / inc_cls_num), like the reference's merged loader — whose shuffling
    cannot change integer counts."""
    model.eval()
    dev = next((p.device for p in model.parameters()), torch.device("cuda", torch.cuda.current_device()))
    nt = task_idx + 1
    loaders = list(dataloaders)[:nt]
    if testing_per_task:
        meter = EvalMeter(nt, dev)
    else:
        if setting == "task-aware":
            raise NotImplementedError("task-aware evaluation needs testing_per_task=True (trainer.py:693-695)")
        assert init_cls_num is not None and inc_cls_num is not None
        bounds = [0]
        for t in range(nt):
            bounds.append(bounds[-1] + (init_cls_num if t == 0 else inc_cls_num))
        meter = EvalMeter(nt, dev, bounds)
    prev = getattr(model, "_defer_metrics", False)
    model._defer_metrics = True          # `inference` keeps its metric on the device (no .item())
    try:
        for t, loader in enumerate(loaders):
            for batch in loader:
                if setting == "task-aware":
                    pred, _ = model.inference(batch, task_id=t)
                else:
                    pred, _ = model.inference(batch)
                label = batch["label"].to(pred.device, torch.int64, non_blocking=True).contiguous()
                # per-task mode counts `int(acc * batch_size)` per batch (trainer.py:644); the merged mode counts exactly (np.sum(preds == labels), :700)
                meter.update(pred.contiguous(), label, t if testing_per_task else -1, reference_rounding=testing_per_task)
    finally:
        model._defer_metrics = prev
    correct, count = meter.result()
    per_task = [round(int(c) * 100 / int(n), 2) if n > 0 else 0 for c, n in zip(correct, count)]
    avg = round(int(correct.sum()) * 100 / int(count.sum()), 2)
    return {"avg_acc": avg, "per_task_acc": per_task}
