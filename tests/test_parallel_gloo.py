"""CPU, world_size 2, gloo: host-side logic of the data-parallel path — the flat-bucket mean all-reduce, batch sharding, and
that replicas fed different shards stay bit-identical after the (oracle) SGD update on the averaged gradient."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from libcontinual_b200.parallel import allreduce_mean_, assert_replicas_identical, shard_batch_size
        from oracle import port as oport
        assert shard_batch_size(128, world) == 64
        g = torch.Generator().manual_seed(1234)
        params = torch.randn(1001, generator=g)                      # identical on every rank
        grads = torch.randn(1001, generator=torch.Generator().manual_seed(rank))   # different per shard
        all_g = torch.stack([torch.randn(1001, generator=torch.Generator().manual_seed(r)) for r in range(world)])
        allreduce_mean_(grads)
        assert torch.allclose(grads, all_g.mean(0), rtol=0, atol=1e-7)
        new_p, m = oport.sgd_momentum_step(params, grads, None, 0.1, 0.9, 5e-4)
        assert_replicas_identical(new_p, what="parameters")
        ok = True
        try:                                                         # a diverged replica must be detected
            assert_replicas_identical(new_p + rank, what="x")
            ok = (rank == 0)
        except RuntimeError:
            ok = rank != 0
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def _vote_worker(rank, world, port, q):
    """L2P's global-batch vote under data parallelism: per-rank top-k counts, SUM all-reduce of the histogram, the reference's majority rule on the sum
    == the rule applied to the whole batch on one rank (oracle/port.py restates prompt.py:380-401)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        from libcontinual_b200.parallel import allreduce_sum_
        from oracle import port as oport
        rng = np.random.default_rng(77)
        sim = rng.standard_normal((64, 10)).astype(np.float32)          # the global batch: identical on every rank
        topk = 5
        whole = oport.l2p_majority_ids_numpy(sim, topk)
        shard = sim[rank * 32:(rank + 1) * 32]
        idx = np.argsort(-shard, axis=1, kind="stable")[:, :topk]
        hist = torch.from_numpy(np.bincount(idx.reshape(-1), minlength=10).astype(np.int32))
        allreduce_sum_(hist)
        full_hist = np.bincount(np.argsort(-sim, axis=1, kind="stable")[:, :topk].reshape(-1), minlength=10)
        ok = np.array_equal(hist.numpy(), full_hist) and np.array_equal(oport.l2p_majority_from_hist_numpy(hist.numpy(), topk), whole)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_l2p_global_vote_histogram_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_vote_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}


def test_flat_bucket_mean_and_replica_consistency_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict(q.get(timeout=10) for _ in range(2))
    assert res == {0: True, 1: True}
