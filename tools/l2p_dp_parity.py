"""Data-parallel L2P with the GLOBAL-batch vote (`sync_vote=True`) against the single-GPU step at the global batch — run under torchrun on N GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/l2p_dp_parity.py

Every rank builds the same model (bench.py's synthetic state), takes its shard of one fixed batch and runs K captured steps (forward / backward, histogram
all-reduce, gradient all-reduce, clip, Adam); rank 0 then repeats the K steps alone on the whole batch with a fresh model.  Printed: whether the voted prompt
ids agree at every step (integer: must be exact) and the relative L2 distance of the trainable arena after K steps (BF16 GEMM tiling differs with the batch
size: fp tolerance), plus the same comparison with the per-rank vote for contrast."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench

GLOBAL_B, K = 32, 3


def build(device, sync_vote):
    from libcontinual_b200.model.l2p import L2P, vit_pt_imnet
    from libcontinual_b200.optim import Adam
    p, prm, key, fc_w, fc_b = bench.l2p_synth_state()
    bb = vit_pt_imnet(pretrained=False, state=p, device=device)
    m = L2P(bb, device, init_cls_num=10, inc_cls_num=10, num_class=100, task_num=10, feat_dim=768, prompt_length=5, pool_size=10, top_k=5,
            pull_constraint_coeff=1.0, sync_vote=sync_vote)
    with torch.no_grad():
        bb.prompt.prompt.copy_(prm)
        # zero-mean keys: U(0, 1) keys share a large mean direction, every sample then ranks the prompts alike and a per-rank vote could not differ from the
        # global one; with these the shards' histograms differ
        bb.prompt.prompt_key.copy_(torch.randn(key.shape, generator=torch.Generator().manual_seed(5)))
        m.network.classifier.weight.copy_(fc_w); m.network.classifier.bias.copy_(fc_b)
    m.after_task(0, None, None, None); m.before_task(1, None, None, None)
    m.train()
    opt = Adam(m.get_parameters(None), lr=0.001875, betas=(0.9, 0.999), weight_decay=0, model=m)
    return m, opt


def run(m, opt, batches, per):
    from libcontinual_b200.trainer import GraphedL2PStep
    step = GraphedL2PStep(m, opt, per)
    ids = []
    for x, y in batches:
        step.run(x, y)
        ids.append(m.ids.clone())
    torch.cuda.synchronize()
    return torch.stack(ids).cpu(), m.theta.detach().clone().cpu(), float(step.loss())


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=device)
    per = GLOBAL_B // world
    full = bench.l2p_batches(K, GLOBAL_B, 10, 20, seed=11)
    shard = [(x[rank * per:(rank + 1) * per].to(device), y[rank * per:(rank + 1) * per].to(device)) for x, y in full]
    out = {}
    for name, sv in (("global_vote", True), ("per_rank_vote", False)):
        m, opt = build(device, sv)
        out[name] = run(m, opt, shard, per)
        del m, opt
        torch.cuda.empty_cache()
    dist.barrier()
    if rank == 0:
        # the single-GPU reference at the global batch: a fresh, non-distributed step (world size 1 process group semantics are the model's own: the
        # step object sees the initialised group, so the reference runs through the plugin's eager order instead)
        m, opt = build(device, False)
        ids1 = []
        for x, y in full:
            opt.zero_grad()
            pred, acc, loss = m.observe({"image": x.to(device), "label": y.to(device)})
            opt.step()
            ids1.append(m.ids.clone())
        torch.cuda.synchronize()
        ids1 = torch.stack(ids1).cpu()
        th1 = m.theta.detach().clone().cpu()
        res = {"world": world, "global_batch": GLOBAL_B, "steps": K}
        res["single_gpu_ids"] = ids1.tolist()
        res["single_gpu_final_loss"] = float(loss)
        for name in out:
            ids, th, shard_loss = out[name]
            res[name] = {"ids_equal_to_single_gpu_every_step": bool(torch.equal(ids, ids1)), "ids": ids.tolist(),
                         "theta_rel_l2_vs_single_gpu": float((th - th1).norm() / th1.norm()), "final_loss_rank0_shard": shard_loss}
        print(json.dumps(res))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
