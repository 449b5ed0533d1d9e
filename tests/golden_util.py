"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle (GPU) tests.  Inputs and initial weights are
regenerated from numpy PCG64 seeds exactly as `oracle/make_golden.py` did; only reference OUTPUTS live in the
.npz fixtures."""
import os

import numpy as np
import torch

from oracle import port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def synth_resnet_state(seed, n_head, feat=64):
    rng = np.random.default_rng(seed)
    p, b = port.cifar_resnet_init(rng)
    bound = 1.0 / np.sqrt(feat)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (n_head, feat)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (n_head,)).astype(np.float32))
    return p, b, fc_w, fc_b


def synth_resnet18_state(seed, n_head):
    rng = np.random.default_rng(seed)
    p, b = port.resnet18_init(rng)
    bound = 1.0 / np.sqrt(512)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (n_head, 512)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (n_head,)).astype(np.float32))
    return p, b, fc_w, fc_b


def synth_alexnet_state(seed, n_tasks=3, cls_per_task=10):
    rng = np.random.default_rng(seed)
    p = port.alexnet_init(rng)
    b = 1.0 / np.sqrt(2048)
    heads = [torch.from_numpy(rng.uniform(-b, b, (cls_per_task, 2048)).astype(np.float32)) for _ in range(n_tasks)]
    return p, heads


def synth_batch(seed, B, lo, hi, img=32):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.standard_normal((B, 3, img, img)).astype(np.float32))
    y = torch.from_numpy(rng.integers(lo, hi, (B,)).astype(np.int64))
    return x, y


def check_summary(g, prefix, named, rtol=1e-4, atol=1e-6, full_rtol=None):
    """Compare a dict of tensors with a fixture summary written by make_golden.summarize()."""
    names = [str(n) for n in g[prefix + "/names"]]
    norms = g[prefix + "/norm"]
    for i, n in enumerate(names):
        t = named[n].detach().double().cpu()
        assert abs(float(t.norm()) - norms[i]) <= rtol * norms[i] + atol, (prefix, n, float(t.norm()), norms[i])
        key = prefix + "/full/" + n
        if key in g.files:
            ref = torch.from_numpy(g[key]).double()
            scale = float(ref.abs().max()) + 1e-12
            err = float((t - ref).abs().max()) / scale
            assert err <= (full_rtol or rtol) , (prefix, n, err)


def synth_vit_state(seed, total_cls=100, pool=10, length=5, depth=12):
    rng = np.random.default_rng(seed)
    p = port.vit_init(rng, depth=depth)
    prm = torch.from_numpy(rng.uniform(0, 1, (1, pool, length, 768)).astype(np.float32))
    key = torch.from_numpy(rng.uniform(0, 1, (pool, 768)).astype(np.float32))
    bound = 1.0 / np.sqrt(768)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (total_cls, 768)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (total_cls,)).astype(np.float32))
    return p, prm, key, fc_w, fc_b


def synth_images(seed, B, lo, hi):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.uniform(0, 1, (B, 3, 224, 224)).astype(np.float32))
    y = torch.from_numpy(rng.integers(lo, hi, (B,)).astype(np.int64))
    return x, y


def l2p_oracle_step(p, prm, key, fc_w, fc_b, x, y, lo, hi, top_k=5, coeff=1.0, gemm_mode="fp32", clip=1.0):
    """One L2P observe() through the oracle (l2p.py:84-107): returns a dict with loss, logits, feat, major ids and the clipped grads."""
    oprm = prm.clone().requires_grad_(True); okey = key.clone().requires_grad_(True)
    ow = fc_w.clone().requires_grad_(True); ob = fc_b.clone().requires_grad_(True)
    feat, rs, major, cls_f = port.l2p_forward(p, oprm, okey, x, top_k, gemm_mode=gemm_mode)
    logits = port.linear_head(feat, ow, ob)
    loss, masked = port.l2p_loss(logits, y, lo, hi, rs, coeff)
    loss.backward()
    if clip is not None:
        torch.nn.utils.clip_grad_norm_([oprm, okey, ow, ob], clip)
    return {"loss": loss.detach(), "logits": logits.detach(), "feat": feat.detach(), "major": major, "cls_features": cls_f, "reduce_sim": rs.detach(),
            "dprompt": oprm.grad, "dkey": okey.grad, "dW": ow.grad, "db": ob.grad}


def synth_lora_state(seed, depth=12, rank=10, n_head=20, slabs="kv"):
    """Same draws as oracle/make_golden.py::synth_lora_state."""
    rng = np.random.default_rng(seed)
    lora = []
    for _ in range(depth):
        d = {}
        for sn in slabs:
            d[f"A_{sn}"] = torch.from_numpy((rng.standard_normal((rank, 768)) / np.sqrt(768 * 3)).astype(np.float32))
            d[f"B_{sn}"] = torch.from_numpy((0.05 * rng.standard_normal((768, rank))).astype(np.float32))
        lora.append(d)
    bound = 1.0 / np.sqrt(768)
    hw = torch.from_numpy(rng.uniform(-bound, bound, (n_head, 768)).astype(np.float32))
    hb = torch.from_numpy(rng.uniform(-bound, bound, (n_head,)).astype(np.float32))
    return lora, hw, hb


def synth_dual_pool(seed, num_class=100):
    """Same draws as oracle/make_golden.py::synth_dual_pool."""
    rng = np.random.default_rng(seed)
    pool = {}
    for l in (0, 1):
        pool[f"g_p_{l}"] = torch.from_numpy(rng.uniform(0, 1, (6, 768)).astype(np.float32))
    for l in (2, 3, 4):
        pool[f"e_p_{l}"] = torch.from_numpy(rng.uniform(0, 1, (10, 20, 768)).astype(np.float32))
        pool[f"e_k_{l}"] = torch.from_numpy(rng.uniform(0, 1, (10, 768)).astype(np.float32))
    bound = 1.0 / np.sqrt(768)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (num_class, 768)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (num_class,)).astype(np.float32))
    return pool, fc_w, fc_b


def synth_coda_pool(seed, num_class=100):
    """Same draws as oracle/make_golden.py::synth_coda_pool."""
    rng = np.random.default_rng(seed)
    pool = {}
    for l in range(5):
        pool[f"e_p_{l}"] = torch.from_numpy((rng.standard_normal((100, 8, 768)) / np.sqrt(768 * 8)).astype(np.float32))
        pool[f"e_k_{l}"] = torch.from_numpy((rng.standard_normal((100, 768)) / np.sqrt(768)).astype(np.float32))
        pool[f"e_a_{l}"] = torch.from_numpy((rng.standard_normal((100, 768)) / np.sqrt(768)).astype(np.float32))
    bound = 1.0 / np.sqrt(768)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (num_class, 768)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (num_class,)).astype(np.float32))
    return pool, fc_w, fc_b


def synth_sdlora_state(seed, n_adapters, n_cls, depth=12, rank=10):
    """Same draws as oracle/make_golden.py::synth_sdlora_state."""
    rng = np.random.default_rng(seed)
    blocks = []
    for _ in range(depth):
        ads = []
        for _ in range(n_adapters):
            d = {}
            for sn in "qv":
                d[f"A_{sn}"] = torch.from_numpy((rng.uniform(-1, 1, (rank, 768)) / np.sqrt(768)).astype(np.float32))
                d[f"B_{sn}"] = torch.from_numpy((0.05 * rng.standard_normal((768, rank))).astype(np.float32))
            ads.append(d)
        blocks.append(ads)
    mags = torch.from_numpy(rng.uniform(0.6, 1.4, (n_adapters,)).astype(np.float32))
    bound = np.sqrt(3.0 / 768)
    hw = torch.from_numpy(rng.uniform(-bound, bound, (n_cls, 768)).astype(np.float32))
    hb = torch.from_numpy(rng.uniform(-0.05, 0.05, (n_cls,)).astype(np.float32))
    return blocks, mags, hw, hb


def sdlora_task_states(upto):
    """The adapter stacks of tasks 0..upto exactly as the golden generator composed them (earlier tasks' adapters carried over)."""
    prev, states = None, []
    for task in range(upto + 1):
        blocks, mags, hw, hb = synth_sdlora_state(970 + task, task + 1, 10 * (task + 1))
        if task > 0:
            for l in range(12):
                blocks[l][:task] = prev[l]
        prev = blocks
        states.append((blocks, mags, hw, hb))
    return states


def synth_input_matrices(seed, L=12, D=768, decay=0.955):
    """Same draws as oracle/make_golden.py::synth_input_matrices."""
    rng = np.random.default_rng(seed)
    out = []
    for l in range(L):
        Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        ev = (decay + 0.002 * l) ** np.arange(D)
        out.append(((Q * ev) @ Q.T).astype(np.float32))
    return np.stack(out)


def synth_timm_vit_state(seed):
    """(reference-named state, timm-named state) of the same synthetic ViT-B/16 weights (oracle/make_golden.py::synth_timm_vit_state)."""
    p = port.vit_init(np.random.default_rng(seed))
    out = {}
    for k, v in p.items():
        out[k.replace("transformer.blocks.", "blocks.").replace(".ln_1.", ".norm1.").replace(".ln_2.", ".norm2.")] = v
    return p, out


def synth_stacked_adapters(seed, n_tasks, depth=12, rank=10):
    """Same draws as oracle/make_golden.py::synth_stacked_adapters."""
    rng = np.random.default_rng(seed)
    blocks = []
    for _ in range(depth):
        ads = []
        for _ in range(n_tasks):
            ads.append({"A_k": torch.from_numpy((rng.standard_normal((rank, 768)) / np.sqrt(768 * 3)).astype(np.float32)),
                        "B_k": torch.from_numpy((0.05 * rng.standard_normal((768, rank))).astype(np.float32)),
                        "A_v": torch.from_numpy((rng.standard_normal((rank, 768)) / np.sqrt(768 * 3)).astype(np.float32)),
                        "B_v": torch.from_numpy((0.05 * rng.standard_normal((768, rank))).astype(np.float32))})
        blocks.append(ads)
    bound = 1.0 / np.sqrt(768)
    hw = torch.from_numpy(rng.uniform(-bound, bound, (n_tasks, 10, 768)).astype(np.float32))
    hb = torch.from_numpy(rng.uniform(-bound, bound, (n_tasks, 10)).astype(np.float32))
    return blocks, hw, hb
