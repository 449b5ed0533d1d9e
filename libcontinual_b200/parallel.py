"""Data-parallel plumbing: one process per GPU, every rank a full replica, ONE exchange step per training
step — the average of the flat gradient arena.  The reference's own DDP path is dead code (core/trainer.py:37-38, :573)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_batch_size(global_batch: int, world_size: int) -> int:
    """Reference semantic for a sharded loader: `batch_size // n_gpu` (core/trainer.py:238)."""
    if global_batch % world_size:
        raise ValueError(f"global batch {global_batch} is not divisible by world size {world_size}")
    return global_batch // world_size


def allreduce_mean_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place mean over ranks of one flat bucket.  NCCL: a single `ncclAvg` all-reduce over NVLink/NVSwitch; other
    backends (gloo in the CPU tests): sum then divide."""
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


def allreduce_sum_(t: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM over ranks of a small integer / float tensor (L2P's prompt histogram: the batch-wide vote of the GLOBAL batch)."""
    if dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t


def assert_replicas_identical(flat: torch.Tensor, group=None, what: str = "state") -> None:
    """Debug check that replicated continual-learning state (theta*, Fisher, teacher, ...) is bit-identical across ranks."""
    world = dist.get_world_size(group)
    if world == 1:
        return
    ref = flat.clone()
    dist.broadcast(ref, src=0, group=group)
    if not torch.equal(ref, flat):
        raise RuntimeError(f"replicated {what} diverged on rank {dist.get_rank(group)}")
