"""TEST INFRASTRUCTURE ONLY.  Generates `tests/golden/*.npz` by running the REAL reference classes
(imported from /root/reference through `oracle/ref_shim.py`) on seeded synthetic inputs, and checks the
oracle restatement (`oracle/port.py`) against them while doing so.

Run here (build container, CPU):   python oracle/make_golden.py
The reference tree does not exist on the GPU box; tests only read the committed .npz files.

Everything random comes from `numpy.random.default_rng(seed)` (PCG64 — platform-stable), never from the
torch RNG, so the tests can regenerate the inputs and initial weights bit-exactly from the seed alone.
"""
from __future__ import annotations

import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import port  # noqa: E402
from oracle.ref_shim import import_reference  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
KEEP_FULL = ("conv_1_3x3.weight", "stage_1.0.conv_a.weight", "stage_2.0.conv_a.weight", "stage_2.0.downsample.0.weight",
             "stage_3.4.conv_b.weight")


# ---- shared synthetic-input helpers (tests import these too) -----------------------------------
def synth_resnet_state(seed: int, n_head: int, feat: int = 64):
    rng = np.random.default_rng(seed)
    p, b = port.cifar_resnet_init(rng)
    bound = 1.0 / np.sqrt(feat)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (n_head, feat)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (n_head,)).astype(np.float32))
    return p, b, fc_w, fc_b


def synth_batch(seed: int, B: int, lo: int, hi: int, img: int = 32):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.standard_normal((B, 3, img, img)).astype(np.float32))
    y = torch.from_numpy(rng.integers(lo, hi, (B,)).astype(np.int64))
    return x, y


class FakeLoader:
    """Stands in for the DataLoader that `EWC.getFisher` iterates (ewc.py:185-202): needs iteration,
    `len()` and `.batch_size` only."""

    def __init__(self, batches, batch_size):
        self.batches, self.batch_size = batches, batch_size

    def __iter__(self):
        return iter({"image": x, "label": y} for x, y in self.batches)

    def __len__(self):
        return len(self.batches)


def summarize(prefix, named, out):
    """Per-tensor (sum, L2 norm) for every tensor + full copies of a few."""
    names = list(named.keys())
    out[prefix + "/names"] = np.array(names)
    out[prefix + "/sum"] = np.array([float(named[n].double().sum()) for n in names])
    out[prefix + "/norm"] = np.array([float(named[n].double().norm()) for n in names])
    for n in names:
        short = n.replace("backbone.", "").replace("network.", "")
        if short in KEEP_FULL or "bn" in short or "classifier" in n or "downsample.1" in short:
            out[prefix + "/full/" + n] = named[n].detach().numpy().copy()


def close(a, b, rtol, atol, what):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item() if a.numel() else 0.0
    ok = torch.allclose(a, b, rtol=rtol, atol=atol)
    print(f"   [{'ok' if ok else 'MISMATCH'}] {what}: max|d|={err:.3e}")
    assert ok, what


def load_ref_backbone(bb, p, b):
    sd = {**p, **b}
    missing = bb.load_state_dict(sd, strict=True)
    return missing


# ---- EWC ---------------------------------------------------------------------------------------
def golden_ewc(core):
    import core.model as M
    print("EWC / cifar_resnet32")
    B, init_cls, inc_cls, lamda = 8, 10, 10, 1000.0
    p, b, fc_w, fc_b = synth_resnet_state(101, 20)
    out = {}

    bb = M.cifar_resnet32()
    load_ref_backbone(bb, p, b)
    ref = M.EWC(bb, 64, 100, device=torch.device("cpu"), init_cls_num=init_cls, inc_cls_num=inc_cls, lamda=lamda)
    ref.before_task(0, None, None, None)
    with torch.no_grad():
        ref.network.classifier.weight.copy_(fc_w[:10]); ref.network.classifier.bias.copy_(fc_b[:10])
    ref.train()
    opt = torch.optim.SGD(ref.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4)

    orc = port.ResNetMethodOracle("ewc", p, b, fc_w[:10], fc_b[:10], init_cls=init_cls, inc_cls=inc_cls, lamda=lamda)

    def ref_step(x, y, tag):
        pred, acc, loss = ref.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward()
        grads = {n: q.grad.clone() for n, q in ref.network.named_parameters()}
        opt.step()
        out[tag + "/loss"] = np.float64(loss.item()); out[tag + "/pred"] = pred.numpy().copy(); out[tag + "/acc"] = np.float64(acc)
        summarize(tag + "/grad", grads, out)
        return pred, acc, loss.detach(), grads

    def both(x, y, tag):
        pr, ar, lr_, gr = ref_step(x, y, tag)
        po, ao, lo, go = orc.step(x, y)
        close(lo, lr_, 1e-5, 1e-6, tag + " loss")
        assert torch.equal(po, pr) and ao == ar
        for n in gr:
            close(go[n], gr[n], 1e-4, 1e-6, tag + " grad " + n) if n.endswith(KEEP_FULL) or "classifier" in n else None
        worst = max(float((go[n] - gr[n]).abs().max() / (gr[n].abs().max() + 1e-12)) for n in gr)
        print(f"   worst rel grad err over all tensors: {worst:.3e}")
        assert worst < 1e-3

    # task 0: two steps
    for s in range(2):
        x, y = synth_batch(1000 + s, B, 0, 10)
        both(x, y, f"t0s{s}")
    summarize("t0/param", dict(ref.network.named_parameters()), out)
    # task boundary: Fisher over 3 batches (last ragged), in train() mode
    fb = [synth_batch(1100 + i, B if i < 2 else 5, 0, 10) for i in range(3)]
    ref.after_task(0, None, FakeLoader(fb, B), None)
    orc.ewc_after_task(fb, B)
    summarize("t0/fisher", ref.fisher, out)
    for n in ref.fisher:
        close(orc.fisher[n], ref.fisher[n], 2e-4, 1e-9, "fisher " + n) if n.endswith(KEEP_FULL) or "classifier" in n else None
    # task 1
    ref.before_task(1, None, None, None)
    with torch.no_grad():
        ref.network.classifier.weight[10:].copy_(fc_w[10:20]); ref.network.classifier.bias[10:].copy_(fc_b[10:20])
    opt = torch.optim.SGD(ref.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4)
    ref.train()
    orc.task_idx = 1
    orc.grow_head(fc_w[:20], fc_b[:20]); orc.reset_optimizer()
    for s in range(3):
        x, y = synth_batch(1200 + s, B, 10, 20)
        both(x, y, f"t1s{s}")
    summarize("t1/param", dict(ref.network.named_parameters()), out)
    summarize("t1/bnbuf", {n: v.float() for n, v in ref.network.named_buffers() if "num_batches" not in n}, out)
    # second boundary exercises the alpha merge with a grown head (ewc.py:129-131)
    fb = [synth_batch(1300 + i, B, 10, 20) for i in range(2)]
    ref.after_task(1, None, FakeLoader(fb, B), None)
    orc.ewc_after_task(fb, B)
    summarize("t1/fisher", ref.fisher, out)
    close(orc.fisher["classifier.weight"], ref.fisher["classifier.weight"], 2e-4, 1e-9, "fisher merge classifier.weight")
    np.savez_compressed(os.path.join(OUT, "ewc_resnet32.npz"), **out)


# ---- iCaRL -------------------------------------------------------------------------------------
def golden_icarl(core):
    import copy
    import core.model as M
    print("iCaRL / cifar_resnet32")
    B, init_cls, inc_cls = 8, 10, 5
    p, b, fc_w, fc_b = synth_resnet_state(202, 100)
    out = {}
    bb = M.cifar_resnet32()
    load_ref_backbone(bb, p, b)
    ref = M.ICarl(bb, 64, 100, device=torch.device("cpu"), init_cls_num=init_cls, inc_cls_num=inc_cls, task_num=11)
    with torch.no_grad():
        ref.network.classifier.weight.copy_(fc_w); ref.network.classifier.bias.copy_(fc_b)
    ref.before_task(0, None, None, None)
    ref.train()
    opt = torch.optim.SGD(ref.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4)
    orc = port.ResNetMethodOracle("icarl", p, b, fc_w, fc_b, init_cls=init_cls, inc_cls=inc_cls)

    def both(x, y, tag):
        pred, acc, loss = ref.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward()
        grads = {n: q.grad.clone() for n, q in ref.network.named_parameters()}
        opt.step()
        out[tag + "/loss"] = np.float64(loss.item()); out[tag + "/pred"] = pred.numpy().copy(); out[tag + "/acc"] = np.float64(acc)
        summarize(tag + "/grad", grads, out)
        po, ao, lo, go = orc.step(x, y)
        close(lo, loss.detach(), 1e-5, 1e-6, tag + " loss")
        assert torch.equal(po, pred) and ao == acc
        worst = max(float((go[n] - grads[n]).abs().max() / (grads[n].abs().max() + 1e-12)) for n in grads)
        print(f"   worst rel grad err over all tensors: {worst:.3e}")
        assert worst < 1e-3

    for s in range(2):
        x, y = synth_batch(2000 + s, B, 0, 10)
        both(x, y, f"t0s{s}")
    # after_task without the disk-backed buffer: the teacher snapshot lines of icarl.py:172-176 only
    ref.old_network = copy.deepcopy(ref.network); ref.old_network.eval()
    ref.prev_cls_num = ref.accu_cls_num; ref.cur_task_id += 1
    ref.before_task(1, None, None, None)
    opt = torch.optim.SGD(ref.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4)
    orc.snapshot_teacher(); orc.prev_cls = 10; orc.accu_cls = 15; orc.task_idx = 1; orc.reset_optimizer()
    for s in range(2):
        x, y = synth_batch(2100 + s, B, 0, 15)     # new classes mixed with exemplars of old ones
        both(x, y, f"t1s{s}")
    summarize("t1/param", dict(ref.network.named_parameters()), out)
    np.savez_compressed(os.path.join(OUT, "icarl_resnet32.npz"), **out)


# ---- LwF ---------------------------------------------------------------------------------------
def golden_lwf(core):
    import core.model as M
    print("LwF / cifar_resnet32 backbone")
    B, init_cls, inc_cls = 8, 10, 10
    p, b, fc_w, fc_b = synth_resnet_state(303, 20)
    out = {}
    bb = M.cifar_resnet32()
    load_ref_backbone(bb, p, b)
    ref = M.LWF(bb, 64, 100, device=torch.device("cpu"), init_cls_num=init_cls, inc_cls_num=inc_cls)
    ref.before_task(0, None, None, None)
    with torch.no_grad():
        ref.classifier.weight.copy_(fc_w[:10]); ref.classifier.bias.copy_(fc_b[:10])
    ref.train()

    def params():
        return list(ref.backbone.parameters()) + list(ref.classifier.parameters())

    def named():
        d = {"backbone." + n: q for n, q in ref.backbone.named_parameters()}
        d.update({"classifier." + n: q for n, q in ref.classifier.named_parameters()})
        return d

    opt = torch.optim.SGD(params(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    orc = port.ResNetMethodOracle("lwf", p, b, fc_w[:10], fc_b[:10], init_cls=init_cls, inc_cls=inc_cls)

    def both(x, y, tag):
        pred, acc, loss = ref.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward()
        grads = {n: q.grad.clone() for n, q in named().items()}
        opt.step()
        out[tag + "/loss"] = np.float64(loss.item()); out[tag + "/pred"] = pred.numpy().copy()
        summarize(tag + "/grad", grads, out)
        po, ao, lo, go = orc.step(x, y)
        close(lo, loss.detach(), 1e-5, 1e-6, tag + " loss")
        assert torch.equal(po, pred)
        worst = max(float((go[n] - grads[n]).abs().max() / (grads[n].abs().max() + 1e-12)) for n in grads)
        print(f"   worst rel grad err over all tensors: {worst:.3e}")
        assert worst < 1e-3

    x, y = synth_batch(3000, B, 0, 10)
    both(x, y, "t0s0")
    ref.before_task(1, None, None, None)          # update_fc + frozen deepcopy of the backbone (lwf.py:44-50)
    with torch.no_grad():
        ref.classifier.weight[10:].copy_(fc_w[10:20]); ref.classifier.bias[10:].copy_(fc_b[10:20])
    ref.train(); ref.old_backbone.eval(); ref.old_fc.eval()
    opt = torch.optim.SGD(params(), lr=0.1, momentum=0.9, weight_decay=5e-4)
    orc.snapshot_teacher(); orc.prev_cls = 10; orc.task_idx = 1
    orc.grow_head(fc_w[:20], fc_b[:20]); orc.reset_optimizer()
    for s in range(2):
        x, y = synth_batch(3100 + s, B, 10, 20)
        both(x, y, f"t1s{s}")
    np.savez_compressed(os.path.join(OUT, "lwf_resnet32.npz"), **out)


def synth_resnet18_state(seed: int, n_head: int):
    rng = np.random.default_rng(seed)
    p, b = port.resnet18_init(rng)
    bound = 1.0 / np.sqrt(512)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (n_head, 512)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (n_head,)).astype(np.float32))
    return p, b, fc_w, fc_b


def golden_lwf18(core):
    """BASELINE config C5's method/backbone pair: the real `LWF` on the real `resnet18` with the tiny-imagenet stem (3x3 conv + maxpool, resnet.py:145-150)
    at 64 x 64: task 0 (CE) and two task-1 steps (CE on the new slice + 3 * KD against the frozen copy)."""
    import core.model as M
    print("LwF / resnet18 (tiny-imagenet stem), 64x64")
    B, init_cls, inc_cls = 4, 10, 10
    p, b, fc_w, fc_b = synth_resnet18_state(1818, 20)
    out = {}
    bb = M.resnet18(args={"dataset": "tiny-imagenet", "init_cls_num": init_cls * 2, "inc_cls_num": inc_cls})
    sd = {**p, **b, "fc.weight": bb.fc.weight.data.clone(), "fc.bias": bb.fc.bias.data.clone()}
    bb.load_state_dict(sd, strict=True)
    ref = M.LWF(bb, 512, 200, device=torch.device("cpu"), init_cls_num=init_cls, inc_cls_num=inc_cls)
    ref.before_task(0, None, None, None)
    with torch.no_grad():
        ref.classifier.weight.copy_(fc_w[:10]); ref.classifier.bias.copy_(fc_b[:10])
    ref.train()

    def named():
        d = {"backbone." + n: q for n, q in ref.backbone.named_parameters() if not n.startswith("fc.")}
        d.update({"classifier." + n: q for n, q in ref.classifier.named_parameters()})
        return d

    opt = torch.optim.SGD(list(named().values()), lr=0.1, momentum=0.9, weight_decay=5e-4)
    orc = port.ResNetMethodOracle("lwf", p, b, fc_w[:10], fc_b[:10], init_cls=init_cls, inc_cls=inc_cls, arch="resnet18", maxpool=True)

    def both(x, y, tag):
        pred, acc, loss = ref.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward()
        assert ref.backbone.fc.weight.grad is None                 # the backbone's own `fc` (resnet.py:191) is never in the forward
        grads = {n: q.grad.clone() for n, q in named().items()}
        opt.step()
        out[tag + "/loss"] = np.float64(loss.item()); out[tag + "/pred"] = pred.numpy().copy()
        summarize(tag + "/grad", grads, out)
        po, ao, lo, go = orc.step(x, y)
        close(lo, loss.detach(), 1e-5, 1e-6, tag + " loss")
        assert torch.equal(po, pred)
        worst = max(float((go[n] - grads[n]).abs().max() / (grads[n].abs().max() + 1e-12)) for n in grads)
        print(f"   worst rel grad err over all tensors: {worst:.3e}")
        assert worst < 1e-3

    x, y = synth_batch(1900, B, 0, 10, img=64)
    both(x, y, "t0s0")
    with torch.no_grad():
        f = ref.backbone(x)
        fo = port.resnet18_forward({k: v.detach() for k, v in orc.p.items()}, orc.b, x, True, True)      # both sides move their running statistics once more
        close(fo["features"], f["features"], 1e-4, 1e-5, "features after one step")
        out["t0s0/features_after"] = f["features"].numpy().copy()
    ref.before_task(1, None, None, None)
    with torch.no_grad():
        ref.classifier.weight[10:].copy_(fc_w[10:20]); ref.classifier.bias[10:].copy_(fc_b[10:20])
    ref.train(); ref.old_backbone.eval(); ref.old_fc.eval()
    opt = torch.optim.SGD(list(named().values()), lr=0.1, momentum=0.9, weight_decay=5e-4)
    orc.snapshot_teacher(); orc.prev_cls = 10; orc.task_idx = 1
    orc.grow_head(fc_w[:20], fc_b[:20]); orc.reset_optimizer()
    for s in range(2):
        x, y = synth_batch(1910 + s, B, 10, 20, img=64)
        both(x, y, f"t1s{s}")
    np.savez_compressed(os.path.join(OUT, "lwf_resnet18.npz"), **out)


# ---- LUCIR --------------------------------------------------------------------------------------
def cifar_to_lucir_name(n: str) -> str:
    """cifar_resnet32 parameter name -> modified_ResNet (resnet32_V2) name: same topology / order, other names."""
    n = n.replace("conv_1_3x3", "conv1").replace("bn_1.", "bn1.")
    for s in (1, 2, 3):
        n = n.replace(f"stage_{s}.", f"layer{s}.")
    return n.replace("conv_a", "conv1").replace("bn_a", "bn1").replace("conv_b", "conv2").replace("bn_b", "bn2")


def golden_lucir(core):
    import core.model as M
    print("LUCIR / resnet32_V2")
    B, init_cls, inc_cls = 8, 10, 5
    p, b, fc_w, fc_b = synth_resnet_state(404, 15)
    rng = np.random.default_rng(4040)
    out = {}
    bb = M.resnet32_V2()
    bb.load_state_dict({cifar_to_lucir_name(k): v for k, v in {**p, **b}.items()}, strict=True)
    ref = M.LUCIR(bb, 64, 100, device=torch.device("cpu"), init_cls_num=init_cls, inc_cls_num=inc_cls, K=2, lw_mr=1, lamda=5, dist=0.5)
    w0 = fc_w[:10].clone()
    with torch.no_grad():
        ref.network.classifier.weight.copy_(w0); ref.network.classifier.sigma.fill_(1.5)
    ref.before_task(0, None, None, None)
    ref.train()
    x, y = synth_batch(4100, B, 0, 10)
    pred, acc, loss = ref.observe({"image": x, "label": y})
    grads = torch.autograd.grad(loss, list(ref.network.parameters()))
    names = [n for n, _ in ref.network.named_parameters()]
    out["t0/loss"] = np.float64(loss.item()); out["t0/pred"] = pred.numpy().copy()
    summarize("t0/grad", dict(zip(names, grads)), out)
    # oracle check (task 0 = plain CE on the cosine logits)
    op = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ob = {k: v.clone() for k, v in b.items()}
    ow, osig = w0.clone().requires_grad_(True), torch.tensor([1.5], requires_grad=True)
    feat = port.cifar_resnet_forward(op, ob, x, True, last_relu=False)["features"]
    l0 = F.cross_entropy(port.cosine_head(feat, ow, osig), y)
    close(l0, loss.detach(), 1e-6, 1e-7, "lucir t0 loss")
    # task 1: reference before_task with the dataset-backed embedding init bypassed (weights set explicitly afterwards)
    ref._init_new_fc = lambda *a, **k: None
    ref.before_task(1, None, None, None)
    w2 = fc_w[10:15].clone()
    with torch.no_grad():
        ref.network.classifier.fc2.weight.copy_(w2)
    ref.train(); ref.ref_model.eval()
    x, y = synth_batch(4101, B, 0, 15)
    y[0], y[1] = 3, 12          # make sure old-class (hard) and new-class samples are both present
    pred, acc, loss = ref.observe({"image": x, "label": y})
    params = dict(ref.network.named_parameters())
    grads = torch.autograd.grad(loss, list(params.values()))
    out["t1/loss"] = np.float64(loss.item()); out["t1/pred"] = pred.numpy().copy(); out["t1/y"] = y.numpy().copy()
    out["t1/cur_lamda"] = np.float64(ref.cur_lamda)
    summarize("t1/grad", dict(zip(params.keys(), grads)), out)
    for n, g in zip(params.keys(), grads):
        if "classifier" in n:
            out["t1/full/" + n] = g.numpy().copy()
    # oracle check of the task-1 loss and head gradients
    W = torch.cat([w0, w2]).requires_grad_(True)
    with torch.no_grad():      # frozen copy taken at before_task(1): its BN statistics are those BEFORE this step's update
        rfeat = port.cifar_resnet_forward({k: v.detach() for k, v in p.items()}, {k: v.clone() for k, v in ob.items()}, x, False, last_relu=False)["features"]
    feat = port.cifar_resnet_forward(op, ob, x, True, last_relu=False)["features"]
    scores = port.cosine_head(feat, W, None)
    l1 = port.lucir_loss(feat, rfeat, osig * scores, scores, y, 10, ref.cur_lamda, 2, 0.5, 1)
    close(l1, loss.detach(), 1e-5, 1e-6, "lucir t1 loss")
    gW, = torch.autograd.grad(l1, [W])
    close(gW[:10], dict(zip(params.keys(), grads))["classifier.fc1.weight"], 1e-4, 1e-7, "lucir d fc1")
    close(gW[10:], dict(zip(params.keys(), grads))["classifier.fc2.weight"], 1e-4, 1e-7, "lucir d fc2")
    np.savez_compressed(os.path.join(OUT, "lucir_resnet32.npz"), **out)


# ---- herding (exemplar selection) ------------------------------------------------------------------
def herding_inputs(seed=2121, ncls=4, n_per=90, D=64):
    rng = np.random.default_rng(seed)
    sizes = [n_per - 7 * (c % 3) for c in range(ncls)]
    feats = np.abs(rng.standard_normal((sum(sizes), D))).astype(np.float32)
    labels = np.concatenate([np.full(s, c, dtype=np.int64) for c, s in enumerate(sizes)])
    return torch.from_numpy(feats), torch.from_numpy(labels)


def golden_herding(core):
    """Drives the REAL `LinearHerdingBuffer.herding_select` with an identity 'backbone' over a tensor-backed dataset, so that the
    recorded index list is the reference's own."""
    from core.model.buffer.linearherdingbuffer import LinearHerdingBuffer
    print("herding_select (LinearHerdingBuffer)")
    raw, labels = herding_inputs()

    class DS(torch.utils.data.Dataset):
        def __init__(self):
            self.images = list(range(raw.shape[0])); self.labels = [int(v) for v in labels]; self.trfms = None
        def __getitem__(self, i):
            return {"image": raw[self.images[i]], "label": self.labels[i]}
        def __len__(self):
            return len(self.labels)

    class Loader:
        dataset = DS()

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.backbone = lambda x: {"features": x}

    buf = LinearHerdingBuffer(buffer_size=100, batch_size=32)
    idx = buf.herding_select(Net(), Loader(), None, 0, 4, np.arange(4), torch.device("cpu"))
    feats = raw / raw.norm(dim=1).view(-1, 1)
    mine = port.herding_select(feats, labels, 100 // 4)
    assert [int(i) for i in idx] == mine, "oracle herding differs from the reference"
    print(f"   [ok] {len(mine)} exemplar indices identical")
    np.savez_compressed(os.path.join(OUT, "herding.npz"), idx=np.array(mine, dtype=np.int64))


# ---- small op-level goldens -----------------------------------------------------------------------
def golden_ops(core):
    from core.model.backbone.prompt import L2P as RefL2PPool
    from core.model.backbone.resnet import CosineLinear, SplitCosineLinear
    print("op-level: L2P pool select, cosine heads, KD, GPM project")
    out = {}
    rng = np.random.default_rng(404)
    # L2P pool (prompt.py:346-406) — several batches, including one with exact score ties
    for case, (B, pool, topk, length, D) in enumerate([(16, 10, 5, 5, 768), (128, 10, 5, 5, 768), (7, 10, 5, 5, 64), (4, 6, 2, 3, 32)]):
        mod = RefL2PPool(length=length, prompt_key=True, pool_size=pool, top_k=topk, num_layers=1, embed_dim=D)
        prm = torch.from_numpy(rng.uniform(0, 1, (1, pool, length, D)).astype(np.float32))
        key = torch.from_numpy(rng.uniform(0, 1, (pool, D)).astype(np.float32))
        q = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32))
        if case == 3:
            key[1] = key[0]; key[4] = key[0]       # exact ties in the similarity
        with torch.no_grad():
            mod.prompt.copy_(prm); mod.prompt_key.copy_(key)
        xe = torch.zeros(B, 3, D)
        bp, rs = mod(xe, cls_features=q)
        obp, ors, oid = port.l2p_select(prm, key, q, topk)
        close(obp, bp, 0, 0, f"l2p case{case} prompts"); close(ors, rs, 1e-6, 1e-7, f"l2p case{case} reduce_sim")
        kn = F.normalize(key, dim=-1); qn = F.normalize(q, dim=-1)
        sim_np = (qn @ kn.T).numpy()
        ids_np = port.l2p_majority_ids_numpy(sim_np, topk)
        # `torch.topk` over the integer histogram (prompt.py:387) breaks count ties in an implementation-defined
        # order (CPU partial sort vs CUDA radix select differ), so: the reference's ids must be A valid top-k of the
        # histogram, and must equal the product rule (count desc, id asc) whenever the top-(k+1) counts are distinct.
        hist = np.bincount(sim_np.argsort(axis=1, kind="stable")[:, ::-1][:, :topk].ravel(), minlength=pool) if case != 3 else None
        if hist is not None:
            top_counts = np.sort(hist)[::-1]
            assert sorted(hist[oid.numpy()].tolist(), reverse=True) == top_counts[:topk].tolist()
            strict = len(set(top_counts[: topk + 1].tolist())) == min(topk + 1, pool)
            if strict:
                assert np.array_equal(ids_np, oid.numpy()), (ids_np, oid, hist)
            else:
                assert sorted(hist[ids_np].tolist(), reverse=True) == top_counts[:topk].tolist()
            out[f"l2p{case}/hist"] = hist; out[f"l2p{case}/strict"] = np.int64(strict)
        out[f"l2p{case}/ids_rule"] = ids_np
        out[f"l2p{case}/shape"] = np.array([B, pool, topk, length, D]); out[f"l2p{case}/ids"] = oid.numpy()
        out[f"l2p{case}/reduce_sim"] = np.float64(rs.item()); out[f"l2p{case}/prompt_sum"] = np.float64(bp.double().sum().item())
        out[f"l2p{case}/ties"] = np.int64(case == 3)
        # gradient of the pull term wrt key (drives the only trainables besides the head)
        mod.zero_grad(); (-rs).backward()
        out[f"l2p{case}/dkey"] = mod.prompt_key.grad.numpy().copy()
    # cosine heads (resnet.py:418-463)
    B, D = 16, 64
    feat = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32))
    cl = CosineLinear(D, 10)
    w1 = torch.from_numpy(rng.uniform(-0.125, 0.125, (10, D)).astype(np.float32))
    with torch.no_grad():
        cl.weight.copy_(w1); cl.sigma.fill_(1.7)
    o = cl(feat)
    close(port.cosine_head(feat, w1, torch.tensor([1.7])), o, 1e-6, 1e-7, "CosineLinear")
    out["cos/out"] = o.detach().numpy().copy()
    sc = SplitCosineLinear(D, 10, 5)
    w2 = torch.from_numpy(rng.uniform(-0.125, 0.125, (5, D)).astype(np.float32))
    with torch.no_grad():
        sc.fc1.weight.copy_(w1); sc.fc2.weight.copy_(w2); sc.sigma.fill_(2.5)
    o2 = sc(feat)
    close(port.cosine_head(feat, torch.cat([w1, w2]), torch.tensor([2.5])), o2, 1e-6, 1e-7, "SplitCosineLinear")
    out["cos/split_out"] = o2.detach().numpy().copy()
    np.savez_compressed(os.path.join(OUT, "ops_small.npz"), **out)


def synth_vit_state(seed: int, total_cls: int = 100, pool: int = 10, length: int = 5, depth: int = 12):
    """ViT-B/16 weights + L2P pool (uniform(0,1) like `nn.init.uniform_`, l2p.py:60) + classifier, all from one numpy Generator."""
    rng = np.random.default_rng(seed)
    p = port.vit_init(rng, depth=depth)
    prm = torch.from_numpy(rng.uniform(0, 1, (1, pool, length, 768)).astype(np.float32))
    key = torch.from_numpy(rng.uniform(0, 1, (pool, 768)).astype(np.float32))
    bound = 1.0 / np.sqrt(768)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (total_cls, 768)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (total_cls,)).astype(np.float32))
    return p, prm, key, fc_w, fc_b


def synth_images(seed: int, B: int, lo: int, hi: int):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.uniform(0, 1, (B, 3, 224, 224)).astype(np.float32))        # ToTensor() range, no Normalize in the l2p yaml
    y = torch.from_numpy(rng.integers(lo, hi, (B,)).astype(np.int64))
    return x, y


def golden_l2p(core):
    """The real `core.model.l2p.L2P` on `vit_pt_imnet` (ViTZoo, 12 x 768): observe() on task 0 and on task 1."""
    from core.model.backbone.vit import vit_pt_imnet
    from core.model.l2p import L2P as RefL2P
    print("L2P / ViT-B/16: reference observe() vs oracle")
    out = {}
    p, prm, key, fc_w, fc_b = synth_vit_state(5150)
    bb = vit_pt_imnet(pretrained=False)
    ref = RefL2P(bb, torch.device("cpu"), init_cls_num=10, inc_cls_num=10, num_class=100, task_num=10, feat_dim=768, prompt_length=5,
                 pool_size=10, top_k=5, pull_constraint_coeff=1.0)
    missing = ref.network.backbone.feat.load_state_dict(p, strict=True)
    with torch.no_grad():
        ref.network.backbone.prompt.prompt.copy_(prm); ref.network.backbone.prompt.prompt_key.copy_(key)
        ref.network.classifier.weight.copy_(fc_w); ref.network.classifier.bias.copy_(fc_b)
    for task, (lo, hi) in enumerate([(0, 10), (10, 20)]):
        if task == 1:
            ref.after_task(0, None, None, None); ref.before_task(1, None, None, None)
        x, y = synth_images(600 + task, 4, lo, hi)
        for q in ref.unfrezeed_params:
            q.grad = None
        pred, acc, loss = ref.observe({"image": x, "label": y})
        # oracle (clip_grad_norm_ is applied by observe() itself: l2p.py:104)
        oprm = prm.clone().requires_grad_(True); okey = key.clone().requires_grad_(True)
        ow = fc_w.clone().requires_grad_(True); ob = fc_b.clone().requires_grad_(True)
        feat, rs, major, cls_f = port.l2p_forward(p, oprm, okey, x, 5)
        ologits = port.linear_head(feat, ow, ob)
        oloss, _ = port.l2p_loss(ologits, y, lo, hi, rs, 1.0)
        oloss.backward()
        torch.nn.utils.clip_grad_norm_([oprm, okey, ow, ob], 1.0)
        close(oloss, loss, 1e-6, 1e-7, f"l2p task{task} loss")
        rp = ref.network.backbone.prompt
        close(oprm.grad, rp.prompt.grad, 1e-5, 1e-8, f"l2p task{task} dprompt")
        close(okey.grad, rp.prompt_key.grad, 1e-5, 1e-8, f"l2p task{task} dkey")
        close(ow.grad, ref.network.classifier.weight.grad, 1e-5, 1e-8, f"l2p task{task} dW")
        close(ob.grad, ref.network.classifier.bias.grad, 1e-5, 1e-8, f"l2p task{task} db")
        with torch.no_grad():
            rlogits, _ = ref.network(x, train=False)
        close(ologits, rlogits, 1e-5, 1e-6, f"l2p task{task} logits")
        out[f"t{task}/loss"] = np.float64(loss.item()); out[f"t{task}/logits"] = rlogits.numpy().copy()
        out[f"t{task}/major"] = major.numpy().copy(); out[f"t{task}/cls_features"] = cls_f.numpy().copy()
        out[f"t{task}/feat"] = feat.detach().numpy().copy()
        out[f"t{task}/dprompt"] = rp.prompt.grad.numpy().copy(); out[f"t{task}/dkey"] = rp.prompt_key.grad.numpy().copy()
        out[f"t{task}/dW"] = ref.network.classifier.weight.grad.numpy().copy(); out[f"t{task}/db"] = ref.network.classifier.bias.grad.numpy().copy()
        out[f"t{task}/pred"] = pred.numpy().copy()
    np.savez_compressed(os.path.join(OUT, "l2p_vit.npz"), **out)


def synth_lora_state(seed: int, depth: int = 12, rank: int = 10, n_head: int = 20, slabs: str = "kv"):
    """Adapters for every block (A ~ orthonormal-ish rows / sqrt(3) scale, B small but non-zero so that the merge is exercised) + one head."""
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


def golden_inflora(core):
    """The real `core.model.InfLoRA_opt.InfLoRA_OPT` on `vit_pt_imnet(attn_layer='MultiHeadAttention_LoRA', lora_rank=10)`: observe() on
    task 0 and task 1 with given adapters, the input-matrix pass (`update_input_matrix`) and the task-0 basis (SVD)."""
    from core.model.backbone.vit import vit_pt_imnet
    from core.model.InfLoRA_opt import InfLoRA_OPT as RefInfLoRA
    from core.utils import init_seed
    print("InfLoRA_OPT / ViT-B/16: reference observe() vs oracle")
    init_seed(42, True)                              # sets PYTHONHASHSEED, read by InfLoRA_opt.py:55
    out = {}
    p, _, _, _, _ = synth_vit_state(5150)
    bb = vit_pt_imnet(pretrained=False, attn_layer="MultiHeadAttention_LoRA", lora_rank=10)
    ref = RefInfLoRA(bb, torch.device("cpu"), init_cls_num=20, inc_cls_num=20, task_num=10, lame=1.0, lamb=0.95, embd_dim=768, use_ca=False,
                     dataset="imagenet-r")
    ref._network.backbone.feat.load_state_dict(p, strict=False)
    for task in (0, 1):
        lora, hw, hb = synth_lora_state(880 + task)
        # what before_task does around the loader pass (InfLoRA_opt.py:207-229), with the adapters given instead of derived
        if task == 1:
            ref._known_classes = ref.init_cls_num
        ref._network.update_fc(None)
        for i, mod in enumerate(ref.attention_modules):
            mod.init_param()
            with torch.no_grad():
                mod.lora_A_k.weight.copy_(lora[i]["A_k"]); mod.lora_B_k.weight.copy_(lora[i]["B_k"])
                mod.lora_A_v.weight.copy_(lora[i]["A_v"]); mod.lora_B_v.weight.copy_(lora[i]["B_v"])
        for name, prm in ref._network.named_parameters():
            prm.requires_grad_(f"classifier_pool.{task}." in name or "lora_B" in name)
            prm.grad = None
        head = ref._network.classifier_pool[task]
        with torch.no_grad():
            head.weight.copy_(hw); head.bias.copy_(hb)
        lo = 0 if task == 0 else 20
        x, y = synth_images(700 + task, 4, lo, lo + 20)
        ref._network.train()
        pred, acc, loss = ref.observe({"image": x, "label": y})
        loss.backward()
        # oracle
        ol = [{k: v.clone().requires_grad_(k.startswith("B_")) for k, v in d.items()} for d in lora]
        ow = hw.clone().requires_grad_(True); ob = hb.clone().requires_grad_(True)
        po = p if task == 0 else p1
        ologits = port.inflora_logits(po, ol, ow, ob, x)
        oloss = F.cross_entropy(ologits, y - lo)
        oloss.backward()
        close(oloss, loss, 1e-6, 1e-7, f"inflora task{task} loss")
        close(ow.grad, head.weight.grad, 1e-4, 2e-6, f"inflora task{task} dW")
        close(ob.grad, head.bias.grad, 1e-4, 2e-6, f"inflora task{task} db")
        dBk = torch.stack([m.lora_B_k.weight.grad for m in ref.attention_modules]); dBv = torch.stack([m.lora_B_v.weight.grad for m in ref.attention_modules])
        close(torch.stack([d["B_k"].grad for d in ol]), dBk, 1e-4, 2e-6, f"inflora task{task} dB_k")
        close(torch.stack([d["B_v"].grad for d in ol]), dBv, 1e-4, 2e-6, f"inflora task{task} dB_v")
        with torch.no_grad():
            rlogits = ref._network(x)
        out[f"t{task}/loss"] = np.float64(loss.item()); out[f"t{task}/logits"] = rlogits.numpy().copy(); out[f"t{task}/pred"] = pred.numpy().copy()
        out[f"t{task}/dW"] = head.weight.grad.numpy().copy(); out[f"t{task}/db"] = head.bias.grad.numpy().copy()
        out[f"t{task}/dB_k"] = dBk.numpy().copy(); out[f"t{task}/dB_v"] = dBv.numpy().copy()
        if task == 0:
            # input-matrix pass (two batches) with the adapters applied, then the task-0 basis of two blocks
            xs = [synth_images(710 + j, 3, 0, 20)[0] for j in range(2)]
            for mod in ref.attention_modules:
                mod.reset_input_matrix()
            with torch.no_grad():
                for xb in xs:
                    ref._network.update_input_matrix(x=xb)
            cur = port.inflora_input_matrices(p, lora, xs)
            proj = torch.from_numpy(np.random.default_rng(99).standard_normal((768, 8)).astype(np.float32))
            for i, mod in enumerate(ref.attention_modules):
                close(cur[i], mod.cur_matrix, 1e-4, 1e-6, f"inflora input matrix {i}")
            out["cov/proj"] = torch.stack([m.cur_matrix @ proj for m in ref.attention_modules]).numpy().copy()
            out["cov/trace"] = np.array([float(m.cur_matrix.trace()) for m in ref.attention_modules])
            for i in (0, 11):
                U, S, _ = torch.linalg.svd(ref.attention_modules[i].cur_matrix, full_matrices=False)
                close(port.inflora_init_A(cur[i], 10).abs(), (U[:, :10].T / math.sqrt(3)).abs(), 1e-3, 1e-5, f"inflora basis {i}")
                out[f"cov/S{i}"] = S[:16].numpy().copy()
                out[f"cov/P{i}"] = (U[:, :10] @ U[:, :10].T @ proj).numpy().copy()        # projector onto the basis (sign / rotation free)
            # merge_weight() (after_task, InfLoRA_opt.py:266-267): the merged QKV weights feed task 1
            for mod in ref.attention_modules:
                mod.merge_weight()
            p1 = dict(p)
            for i in range(12):
                p1[f"transformer.blocks.{i}.attn.qkv.weight"] = port.lora_merge_qkv(p[f"transformer.blocks.{i}.attn.qkv.weight"], lora[i]["A_k"], lora[i]["B_k"],
                                                                                   lora[i]["A_v"], lora[i]["B_v"])
                close(p1[f"transformer.blocks.{i}.attn.qkv.weight"], ref.attention_modules[i].qkv.weight, 0, 0, f"merged qkv {i}")
    np.savez_compressed(os.path.join(OUT, "inflora_vit.npz"), **out)


def synth_dual_pool(seed: int, num_class: int = 100):
    """DualPrompt pool (uniform(0,1) like `tensor_prompt`, prompt.py:409-418) + classifier from one numpy Generator."""
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


def ref_prompt_forward(ref_zoo, x, train, task_id):
    """`ViTZoo.forward` prompt branch (vit.py:121-131) + `VisionTransformer.forward` non-l2p branch (transformer.py:2263-2296) driven through the REAL
    reference modules (patch_embed, blocks with prompt=(pk, pv), the prompt pool's forward, norm).  The only change: the pool losses are summed
    out of place — the reference's in-place `prompt_loss += loss` on a leaf tensor raises on CPU under torch >= 2.1."""
    vt = ref_zoo.feat
    with torch.no_grad():
        q, _ = vt(x)
        q = q[:, 0, :]
    B = x.shape[0]
    t = vt.patch_embed(x)
    t = torch.cat((vt.cls_token.expand(B, -1, -1), t), dim=1)
    t = vt.pos_drop(t + vt.pos_embed[:, :t.size(1), :])
    ploss = torch.zeros(())
    for i, blk in enumerate(vt.transformer.blocks):
        p_list, loss, t = ref_zoo.prompt.forward(q, i, t, train=train, task_id=task_id)
        if train:
            ploss = ploss + loss
        t = blk(t.permute(1, 0, 2), register_hook=False, prompt=p_list).permute(1, 0, 2)
    out = vt.norm(t)[:, 0, :]
    return out, ploss, q


def golden_dualprompt(core):
    """The real `core.model.dualprompt.DualPrompt` (pool + ViT blocks + classifier) on task 0 and task 1, and the per-sample selection at inference."""
    from core.model.backbone.vit import vit_pt_imnet
    from core.model.dualprompt import DualPrompt as RefDual
    print("DualPrompt / ViT-B/16: reference modules vs oracle")
    out = {}
    p = synth_vit_state(5150)[0]
    pool, fc_w, fc_b = synth_dual_pool(930)
    bb = vit_pt_imnet(pretrained=False)
    ref = RefDual(bb, 768, 100, device=torch.device("cpu"), task_num=10, init_cls_num=10, inc_cls_num=10, g_prompt_length=6, e_prompt_length=20)
    bb.feat.load_state_dict(p, strict=True)
    rp = bb.prompt
    with torch.no_grad():
        for k, v in pool.items():
            getattr(rp, k).copy_(v)
    for task in (0, 1):
        ref.before_task(task, None, None, None)
        n = ref.network.classifier.out_features
        with torch.no_grad():
            ref.network.classifier.weight.copy_(fc_w[:n]); ref.network.classifier.bias.copy_(fc_b[:n])
        for q_ in ref.get_parameters(None):
            q_.grad = None
        lo = 10 * task
        x, y = synth_images(750 + task, 4, lo, lo + 10)
        # observe() (dualprompt.py:89-104) on the out-of-place forward
        feat, ploss, q = ref_prompt_forward(bb, x, True, task)
        logit = ref.network.classifier(feat)
        logit[:, :ref.last_out_dim] = -float("inf")
        loss = ploss + (ref.loss_fn(logit, y) * ref.dw_k[-1 * torch.ones(y.size()).long()]).mean()
        loss.backward()
        pred = torch.argmax(logit, dim=1)
        # oracle
        op = {k: v.clone().requires_grad_(True) for k, v in pool.items()}
        ow = fc_w[:n].clone().requires_grad_(True); ob = fc_b[:n].clone().requires_grad_(True)
        ofeat, oploss, oq, _ = port.dualprompt_forward(p, op, x, task, True)
        ologits = port.linear_head(ofeat, ow, ob)
        oloss = port.dualprompt_loss(ologits, y, ref.last_out_dim, oploss)
        oloss.backward()
        close(oloss, loss, 1e-5, 1e-6, f"dual task{task} loss")
        close(oq, q, 1e-4, 1e-5, f"dual task{task} query")
        close(ofeat, feat, 1e-4, 1e-5, f"dual task{task} feat")
        for k in pool:
            g_ref = getattr(rp, k).grad
            close(op[k].grad, g_ref, 1e-3, 1e-9, f"dual task{task} d{k}")
            out[f"t{task}/d{k}"] = g_ref.numpy().copy()
        close(ow.grad, ref.network.classifier.weight.grad, 1e-4, 2e-6, f"dual task{task} dW")
        out[f"t{task}/loss"] = np.float64(loss.item()); out[f"t{task}/ploss"] = np.float64(ploss.item()); out[f"t{task}/pred"] = pred.numpy().copy()
        out[f"t{task}/feat"] = feat.detach().numpy().copy(); out[f"t{task}/query"] = q.numpy().copy()
        out[f"t{task}/dW"] = ref.network.classifier.weight.grad.numpy().copy(); out[f"t{task}/db"] = ref.network.classifier.bias.grad.numpy().copy()
        # inference: per-sample top-1 key (prompt.py:290-292) through the real pool + blocks
        with torch.no_grad():
            ifeat, _, _ = ref_prompt_forward(bb, x, False, task)
            ilogits = ref.network.classifier(ifeat)
            ofeat_i, _, _, ids = port.dualprompt_forward(p, pool, x, task, False)
        close(ofeat_i, ifeat, 1e-4, 1e-5, f"dual task{task} inference feat")
        out[f"t{task}/inf_logits"] = ilogits.numpy().copy()
        out[f"t{task}/inf_ids"] = torch.stack([ids[l] for l in (2, 3, 4)]).numpy().copy()
        ref.after_task(task, None, None, None)
    np.savez_compressed(os.path.join(OUT, "dualprompt_vit.npz"), **out)


def synth_coda_pool(seed: int, num_class: int = 100, nk: int = 10):
    """CodaPrompt pool: the first nk components of every layer are what the reference ever uses; random non-degenerate values (unit-scale rows)."""
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


def golden_codaprompt(core):
    """The real `core.model.codaprompt.CodaPrompt` (pool + ViT blocks + classifier) on task 0 and task 1."""
    from core.model.backbone.vit import vit_pt_imnet
    from core.model.codaprompt import CodaPrompt as RefCoda
    print("CodaPrompt / ViT-B/16: reference modules vs oracle")
    out = {}
    p = synth_vit_state(5150)[0]
    pool, fc_w, fc_b = synth_coda_pool(940)
    bb = vit_pt_imnet(pretrained=False)
    torch.manual_seed(7)
    ref = RefCoda(bb, 768, 100, device=torch.device("cpu"), task_num=10, init_cls_num=10, inc_cls_num=10, prompt_length=8, pool_size=100, mu=0.0)
    bb.feat.load_state_dict(p, strict=True)
    rp = bb.prompt
    # the reference pool's own initialisation (uniform + Gram-Schmidt of the first 10 rows, zeros elsewhere) from a known RNG state: recorded for
    # the init-parity test of libcontinual_b200.model.codaprompt.CodaPromptPool
    from core.model.backbone.prompt import CodaPrompt as RefPool
    torch.manual_seed(7)
    ip = RefPool(768, 10, [100, 8, 0.0])
    out["init/e_k_0_gram"] = (ip.e_k_0[:10] @ ip.e_k_0[:10].T).detach().numpy().copy()
    out["init/e_p_3_tail_absmax"] = np.float64(ip.e_p_3[10:].abs().max().item())
    out["init/e_a_2_head"] = ip.e_a_2[:10, :8].detach().numpy().copy()
    out["init/e_p_4_head"] = ip.e_p_4[:10, 3, :8].detach().numpy().copy()
    with torch.no_grad():
        for k, v in pool.items():
            getattr(rp, k).copy_(v)
    assert rp.task_count == 0
    for task in (0, 1):
        ref.before_task(task, None, None, None)
        assert rp.task_count == 0                      # never advanced by the reference (no caller of process_task_count)
        n = ref.network.classifier.out_features
        with torch.no_grad():
            ref.network.classifier.weight.copy_(fc_w[:n]); ref.network.classifier.bias.copy_(fc_b[:n])
        for q_ in ref.get_parameters(None):
            q_.grad = None
        lo = 10 * task
        x, y = synth_images(760 + task, 4, lo, lo + 10)
        feat, ploss, q = ref_prompt_forward(bb, x, True, task)
        logit = ref.network.classifier(feat)
        logit[:, :ref.last_out_dim] = -float("inf")
        loss = ploss + (ref.loss_fn(logit, y) * ref.dw_k[-1 * torch.ones(y.size()).long()]).mean()
        loss.backward()
        pred = torch.argmax(logit, dim=1)
        op = {k: v.clone().requires_grad_(True) for k, v in pool.items()}
        ow = fc_w[:n].clone().requires_grad_(True); ob = fc_b[:n].clone().requires_grad_(True)
        ofeat, oq = port.codaprompt_forward(p, op, x, 10)
        oloss = port.dualprompt_loss(port.linear_head(ofeat, ow, ob), y, ref.last_out_dim, torch.zeros(()))
        oloss.backward()
        close(oloss, loss, 1e-5, 1e-6, f"coda task{task} loss")
        close(ofeat, feat, 1e-4, 1e-5, f"coda task{task} feat")
        for k in pool:
            g_ref = getattr(rp, k).grad
            close(op[k].grad, g_ref, 1e-3, 1e-10, f"coda task{task} d{k}")
            assert float(g_ref[10:].abs().max()) == 0.0
            out[f"t{task}/d{k}"] = g_ref[:10].numpy().copy()
        close(ow.grad, ref.network.classifier.weight.grad, 1e-4, 2e-6, f"coda task{task} dW")
        out[f"t{task}/loss"] = np.float64(loss.item()); out[f"t{task}/pred"] = pred.numpy().copy()
        out[f"t{task}/feat"] = feat.detach().numpy().copy()
        out[f"t{task}/dW"] = ref.network.classifier.weight.grad.numpy().copy(); out[f"t{task}/db"] = ref.network.classifier.bias.grad.numpy().copy()
        with torch.no_grad():
            ifeat, _, _ = ref_prompt_forward(bb, x, False, task)
            out[f"t{task}/inf_logits"] = ref.network.classifier(ifeat).numpy().copy()
        ref.after_task(task, None, None, None)
    np.savez_compressed(os.path.join(OUT, "codaprompt_vit.npz"), **out)


def synth_sdlora_state(seed: int, n_adapters: int, n_cls: int, depth: int = 12, rank: int = 10):
    """Adapters of every block for `n_adapters` tasks (small non-zero B so that every term is exercised), magnitudes, classifier."""
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


def golden_sdlora(core):
    """The real `core.model.sd_lora.SD_LoRA` on `vit_pt_imnet(attn_layer='MultiHeadAttention_SDLoRA', lora_rank=10)`: observe() + backward on task 0 (one
    adapter) and task 2 (three adapters: two frozen + the current one; all three magnitudes train)."""
    from core.model.backbone.vit import vit_pt_imnet
    from core.model.sd_lora import SD_LoRA as RefSD
    print("SD_LoRA / ViT-B/16: reference observe() vs oracle")
    out = {}
    p = synth_vit_state(5150)[0]
    bb = vit_pt_imnet(pretrained=False, attn_layer="MultiHeadAttention_SDLoRA", lora_rank=10)
    ref = RefSD(bb, torch.device("cpu"), init_cls_num=10, inc_cls_num=10, task_num=10, embd_dim=768, init_mag=1.0, rank_reduction=[False, 4, 8, 8, 6],
                knowledge_dist=[False, 9e-4], dataset="cifar100")
    bb.feat.load_state_dict(p, strict=False)
    for task in (0, 1, 2):
        ref.before_task(task, None, None, None)
        nad, ncls = task + 1, 10 * (task + 1)
        blocks, mags, hw, hb = synth_sdlora_state(970 + task, nad, ncls)
        if task > 0:                                   # earlier tasks' adapters stay what they were
            for l in range(12):
                blocks[l][:task] = prev_blocks[l]
        with torch.no_grad():
            for l, mod in enumerate(ref.attention_modules):
                for i in range(nad):
                    mod.lora_A_q_list[i].weight.copy_(blocks[l][i]["A_q"]); mod.lora_B_q_list[i].weight.copy_(blocks[l][i]["B_q"])
                    mod.lora_A_v_list[i].weight.copy_(blocks[l][i]["A_v"]); mod.lora_B_v_list[i].weight.copy_(blocks[l][i]["B_v"])
            for i in range(nad):
                ref.attention_modules[0].mag_lora[i].copy_(mags[i:i + 1])
            ref._network.classifier.weight.copy_(hw); ref._network.classifier.bias.copy_(hb)
        prev_blocks = blocks
        if task == 1:
            ref.after_task(task, None, None, None)
            continue
        for q_ in ref._network.parameters():
            q_.grad = None
        lo = 10 * task
        x, y = synth_images(780 + task, 4, lo, lo + 10)
        ref._network.train()
        pred, acc, loss = ref.observe({"image": x, "label": y})
        loss.backward()
        # oracle
        ob_ = [[{k: v.clone().requires_grad_(i == task) for k, v in ad.items()} for i, ad in enumerate(blocks[l])] for l in range(12)]
        om = [mags[i:i + 1].clone().requires_grad_(True) for i in range(nad)]
        ow = hw.clone().requires_grad_(True); obias = hb.clone().requires_grad_(True)
        ologits = port.sdlora_logits(p, ob_, om, ow, obias, x)
        oloss = F.cross_entropy(ologits[:, lo:], y - lo)
        oloss.backward()
        close(oloss, loss, 1e-5, 1e-6, f"sdlora task{task} loss")
        rm = ref.attention_modules
        for nm, key, lst in (("dA_q", "A_q", "lora_A_q_list"), ("dB_q", "B_q", "lora_B_q_list"), ("dA_v", "A_v", "lora_A_v_list"), ("dB_v", "B_v", "lora_B_v_list")):
            g_ref = torch.stack([getattr(m, lst)[task].weight.grad for m in rm])
            g_orc = torch.stack([ob_[l][task][key].grad for l in range(12)])
            close(g_orc, g_ref, 1e-3, 1e-4 * float(g_ref.abs().max()), f"sdlora task{task} {nm}")
            out[f"t{task}/{nm}"] = g_ref.numpy().copy()
        gm_ref = torch.cat([rm[0].mag_lora[i].grad for i in range(nad)])
        close(torch.cat([m.grad for m in om]), gm_ref, 1e-3, 1e-4 * float(gm_ref.abs().max()), f"sdlora task{task} dmag")
        close(ow.grad, ref._network.classifier.weight.grad, 1e-4, 2e-6, f"sdlora task{task} dW")
        out[f"t{task}/dmag"] = gm_ref.numpy().copy()
        out[f"t{task}/dW"] = ref._network.classifier.weight.grad.numpy().copy(); out[f"t{task}/db"] = ref._network.classifier.bias.grad.numpy().copy()
        with torch.no_grad():
            out[f"t{task}/logits"] = ref._network(x).numpy().copy()
        out[f"t{task}/loss"] = np.float64(loss.item()); out[f"t{task}/pred"] = pred.numpy().copy()
        # frozen adapters of earlier tasks receive no gradient in the reference
        if task > 0:
            assert rm[3].lora_B_q_list[0].weight.grad is None
        ref.after_task(task, None, None, None)
    np.savez_compressed(os.path.join(OUT, "sdlora_vit.npz"), **out)


def synth_input_matrices(seed: int, L: int = 12, D: int = 768, decay: float = 0.955):
    """Synthetic PSD input matrices with a geometric spectrum and random eigenvectors (one per block), reproducible from the seed."""
    rng = np.random.default_rng(seed)
    out = []
    for l in range(L):
        Q, _ = np.linalg.qr(rng.standard_normal((D, D)))
        ev = (decay + 0.002 * l) ** np.arange(D)
        out.append(((Q * ev) @ Q.T).astype(np.float32))
    return np.stack(out)


def golden_dualgpm(core):
    """The real `InfLoRA_OPT._update_feature` (InfLoRA_opt.py:278-362) on preset input matrices (an empty loader leaves `cur_matrix` as given), four
    tasks in a row with thresholds that exercise growth, the 'remove' -> 'retain' swap and the 'retain' shrink; recorded: basis sizes, types and a
    random projection of each projector F F^T."""
    from core.model.backbone.vit import vit_pt_imnet
    from core.model.InfLoRA_opt import InfLoRA_OPT as RefInfLoRA
    from core.utils import init_seed
    print("DualGPM feature update: reference vs libcontinual_b200.model.inflora.dualgpm_update")
    sys.path.insert(0, os.path.dirname(HERE))
    from libcontinual_b200.model.inflora import dualgpm_update
    init_seed(42, True)
    bb = vit_pt_imnet(pretrained=False, attn_layer="MultiHeadAttention_LoRA", lora_rank=10)
    # lamb / lame chosen so that the captured subspace passes half the dimension at task 1 (swap to 'retain') and is shrunk at task 2
    ref = RefInfLoRA(bb, torch.device("cpu"), init_cls_num=20, inc_cls_num=20, task_num=4, lame=0.9999, lamb=0.999, embd_dim=768, use_ca=False,
                     dataset="imagenet-r")
    proj = np.random.default_rng(7).standard_normal((768, 4)).astype(np.float32)
    mine_f, mine_t = [], []
    out = {}
    for task in range(4):
        acts = synth_input_matrices(1200 + task)
        for i, mod in enumerate(ref.attention_modules):
            mod.cur_matrix = torch.from_numpy(acts[i].copy()); mod.n_cur_matrix = 1
        try:
            ref._update_feature(task, [], None)
        except TypeError as e:
            # the 'retain' shrink branch (InfLoRA_opt.py:343-351) mixes a torch Tensor with numpy arrays and raises under numpy 2 / torch 2.11
            # (it is written for numpy 1.x / torch 2.0.1): no reference output to pin there; ours must still run
            print(f"   task {task}: reference raises {type(e).__name__} in the 'retain' branch -> not pinned ({e})")
            dualgpm_update(acts, mine_f, mine_t, task, 4, 0.9999, 0.999)
            out["unpinned_from_task"] = np.int64(task)
            break
        dualgpm_update(acts, mine_f, mine_t, task, 4, 0.9999, 0.999)
        sizes = np.array([f.shape[1] for f in ref.feature_list])
        print(f"   task {task}: sizes {sizes.tolist()} types {sorted(set(ref.project_type))}")
        assert [f.shape[1] for f in mine_f] == sizes.tolist() and mine_t == ref.project_type, (task, [f.shape[1] for f in mine_f], sizes.tolist(), mine_t)
        P_ref = np.stack([(f @ (f.T @ proj)) for f in ref.feature_list])
        P_mine = np.stack([(f @ (f.T @ proj)) for f in mine_f])
        close(P_mine, P_ref, 1e-3, 1e-3, f"dualgpm task{task} projectors")
        out[f"t{task}/sizes"] = sizes; out[f"t{task}/types"] = np.array([t == "retain" for t in ref.project_type]); out[f"t{task}/P"] = P_ref.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "dualgpm.npz"), **out)


def synth_timm_vit_state(seed: int):
    """The synthetic ViT-B/16 weights of `port.vit_init` under timm's state-dict names (vit_inflora.py: blocks.{i}.norm1 / attn / norm2 / mlp)."""
    p = port.vit_init(np.random.default_rng(seed))
    out = {}
    for k, v in p.items():
        k = k.replace("transformer.blocks.", "blocks.").replace(".ln_1.", ".norm1.").replace(".ln_2.", ".norm2.")
        out[k] = v
    return p, out


def synth_stacked_adapters(seed: int, n_tasks: int, depth: int = 12, rank: int = 10):
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


def golden_inflora_orig(core):
    """The real `core.model.InfLoRA.InfLoRA` on the real `ViT_lora_co` (SiNet.py) encoder: observe() + backward on task 0 and task 1 (two stacked
    adapters, only the second trains), the cur-matrix pass, and `update_DualGPM` on preset matrices.  `SiNet_vit.__init__` itself cannot run offline
    (it builds the encoder through timm's pretrained-model factory): the object is assembled from the same parts (ViT_lora_co + the two head pools)."""
    import torch.nn as nn
    from core.model.backbone.SiNet import SiNet_vit, ViT_lora_co
    from core.model.backbone.vit_inflora import Attention_LoRA
    from core.model.InfLoRA import InfLoRA as RefInfLoRA
    sys.path.insert(0, os.path.dirname(HERE))
    from libcontinual_b200.model.inflora_orig import dualgpm_update_v1
    print("InfLoRA (original) / timm-style ViT-B/16: reference observe() vs oracle")
    out = {}
    p, p_timm = synth_timm_vit_state(5150)
    sn = SiNet_vit.__new__(SiNet_vit)
    nn.Module.__init__(sn)
    sn.image_encoder = ViT_lora_co(patch_size=16, embed_dim=768, depth=12, num_heads=12, n_tasks=10, rank=10)
    sn.class_num = 10
    sn.classifier_pool = nn.ModuleList([nn.Linear(768, 10, bias=True) for _ in range(10)])
    sn.classifier_pool_backup = nn.ModuleList([nn.Linear(768, 10, bias=True) for _ in range(10)])
    sn.numtask = 0
    missing = sn.image_encoder.load_state_dict(p_timm, strict=False)
    assert not [k for k in missing.missing_keys if "lora" not in k and "grow" not in k and "head" not in k], missing.missing_keys
    ref = RefInfLoRA(sn, 768, 100, inc_cls_num=10, device=torch.device("cpu"), lame=1.0, lamb=0.6, total_sessions=10)
    mods = [m for m in ref._network.modules() if isinstance(m, Attention_LoRA)]
    blocks, hw, hb = synth_stacked_adapters(990, 2)
    for task in (0, 1):
        # before_task's bookkeeping (InfLoRA.py:112-138) without the loader pass
        ref._known_classes = ref._total_classes
        ref._cur_task += 1
        ref._total_classes = ref._known_classes + ref.inc_cls_num
        ref._network.update_fc(ref._total_classes)
        for name, prm in ref._network.named_parameters():
            prm.requires_grad_(any(f"{s}.{task}." in name for s in ("classifier_pool", "lora_B_k", "lora_B_v")))
            prm.grad = None
        with torch.no_grad():
            for l, mod in enumerate(mods):
                mod.lora_A_k[task].weight.copy_(blocks[l][task]["A_k"]); mod.lora_B_k[task].weight.copy_(blocks[l][task]["B_k"])
                mod.lora_A_v[task].weight.copy_(blocks[l][task]["A_v"]); mod.lora_B_v[task].weight.copy_(blocks[l][task]["B_v"])
            ref._network.classifier_pool[task].weight.copy_(hw[task]); ref._network.classifier_pool[task].bias.copy_(hb[task])
        lo = 10 * task
        x, y = synth_images(800 + task, 4, lo, lo + 10)
        pred, acc, loss = ref.observe({"image": x, "label": y})
        loss.backward()
        ob_ = [[{k: v.clone().requires_grad_(i == task and k.startswith("B_")) for k, v in ad.items()} for i, ad in enumerate(blocks[l][:task + 1])] for l in range(12)]
        ow = hw[task].clone().requires_grad_(True); obias = hb[task].clone().requires_grad_(True)
        ologits = port.inflora_orig_logits(p, ob_, ow, obias, x)
        oloss = F.cross_entropy(ologits, y - lo)
        oloss.backward()
        close(oloss, loss, 1e-5, 1e-6, f"inflora-orig task{task} loss")
        dBk = torch.stack([m.lora_B_k[task].weight.grad for m in mods]); dBv = torch.stack([m.lora_B_v[task].weight.grad for m in mods])
        close(torch.stack([ob_[l][task]["B_k"].grad for l in range(12)]), dBk, 1e-3, 1e-4 * float(dBk.abs().max()), f"inflora-orig task{task} dB_k")
        close(torch.stack([ob_[l][task]["B_v"].grad for l in range(12)]), dBv, 1e-3, 1e-4 * float(dBv.abs().max()), f"inflora-orig task{task} dB_v")
        head = ref._network.classifier_pool[task]
        close(ow.grad, head.weight.grad, 1e-4, 2e-6, f"inflora-orig task{task} dW")
        if task == 1:
            assert mods[2].lora_B_k[0].weight.grad is None
        with torch.no_grad():
            out[f"t{task}/logits"] = ref._network(x)["logits"].numpy().copy()
            out[f"t{task}/interface"] = ref._network.interface(x).numpy().copy()
        out[f"t{task}/loss"] = np.float64(loss.item()); out[f"t{task}/pred"] = pred.numpy().copy()
        out[f"t{task}/dB_k"] = dBk.numpy().copy(); out[f"t{task}/dB_v"] = dBv.numpy().copy()
        out[f"t{task}/dW"] = head.weight.grad.numpy().copy(); out[f"t{task}/db"] = head.bias.grad.numpy().copy()
        if task == 1:
            # cur-matrix pass (two batches) with both adapters applied
            xs = [synth_images(810 + j, 3, 0, 20)[0] for j in range(2)]
            with torch.no_grad():
                for xb in xs:
                    ref._network(xb, get_cur_feat=True)
            proj = torch.from_numpy(np.random.default_rng(99).standard_normal((768, 8)).astype(np.float32))
            out["cov/proj"] = torch.stack([m.cur_matrix @ proj for m in mods]).numpy().copy()
            out["cov/trace"] = np.array([float(m.cur_matrix.trace()) for m in mods])
    # update_DualGPM on preset matrices: four sessions with lame = 1.0, lamb = 0.999 style thresholds as in golden_dualgpm
    ref2 = RefInfLoRA.__new__(RefInfLoRA)
    ref2.feature_list, ref2.project_type, ref2.lame, ref2.lamb, ref2.total_sessions = [], [], 0.9999, 0.999, 4
    mine_f, mine_t = [], []
    projn = np.random.default_rng(7).standard_normal((768, 4)).astype(np.float32)
    layers = [0, 10, 11]
    for task in range(4):
        acts = synth_input_matrices(1200 + task)[layers]
        ref2._cur_task = task
        try:
            ref2.update_DualGPM([a.copy() for a in acts])      # ndarrays: what np.linalg.svd returned for the tensors under the pinned numpy 1.x
        except Exception as e:
            print(f"   session {task}: reference raises {type(e).__name__}: {e} -> not pinned")
            out["gpm/unpinned_from_task"] = np.int64(task)
            break
        dualgpm_update_v1(list(acts), mine_f, mine_t, task, 4, 0.9999, 0.999)
        sizes = [f.shape[1] for f in ref2.feature_list]
        print(f"   session {task}: sizes {sizes} types {ref2.project_type}")
        assert [f.shape[1] for f in mine_f] == sizes and mine_t == ref2.project_type, (task, [f.shape[1] for f in mine_f], sizes, mine_t)
        P_ref = np.stack([f @ (f.T @ projn) for f in ref2.feature_list]); P_mine = np.stack([f @ (f.T @ projn) for f in mine_f])
        close(P_mine, P_ref, 1e-3, 1e-3, f"update_DualGPM session {task} projectors")
        out[f"gpm/t{task}/sizes"] = np.array(sizes); out[f"gpm/t{task}/types"] = np.array([t == "retain" for t in ref2.project_type])
        out[f"gpm/t{task}/P"] = P_ref.astype(np.float32)
    else:
        out["gpm/unpinned_from_task"] = np.int64(4)
    np.savez_compressed(os.path.join(OUT, "inflora_orig_vit.npz"), **out)


def synth_alexnet_state(seed: int, n_tasks: int = 3, cls_per_task: int = 10):
    rng = np.random.default_rng(seed)
    p = port.alexnet_init(rng)
    b = 1.0 / np.sqrt(2048)
    heads = [torch.from_numpy(rng.uniform(-b, b, (cls_per_task, 2048)).astype(np.float32)) for _ in range(n_tasks)]
    return p, heads


def golden_gpm(core):
    """The real `GPM` on the real `AlexNet_TRGP` (gpm.py:43-206, alexnet.py:94-156), network in eval() so that dropout is off (BatchNorm has
    track_running_stats=False and normalises with batch statistics in either mode): task-0 step, basis construction from 125 samples, task-1 step with
    the projected gradients and frozen BN affine, basis growth."""
    import core.model as M
    from core.model.backbone.alexnet import AlexNet_TRGP
    print("GPM / AlexNet_TRGP")
    B = 16
    p, heads = synth_alexnet_state(4040)
    out = {}
    bb = AlexNet_TRGP()
    bb.load_state_dict(p, strict=True)
    ref = M.GPM(bb, torch.device("cpu"), init_cls_num=10, inc_cls_num=10, task_num=3)
    with torch.no_grad():
        for t in range(3):
            ref.network.classifiers[t].weight.copy_(heads[t])
    orc = port.GPMOracle(p, heads, 10, 10, lr=0.01)
    rng = np.random.default_rng(4141)
    pool = torch.from_numpy(rng.standard_normal((160, 3, 32, 32)).astype(np.float32))

    def named_grads():
        d = {n: q.grad.clone() for n, q in ref.network.named_parameters() if q.grad is not None}
        return {k.replace("backbone.", ""): v for k, v in d.items()}

    def both(x, y, tag):
        opt = torch.optim.SGD([q for q in ref.network.parameters()], lr=0.01)
        ref.network.eval()
        opt.zero_grad()
        pred, acc, loss = ref.observe({"image": x, "label": y})
        grads = named_grads()
        opt.step()
        out[tag + "/loss"] = np.float64(loss.item()); out[tag + "/pred"] = pred.numpy().copy()
        summarize(tag + "/grad", grads, out)
        po, ao, lo, go = orc.step(x, y)
        close(lo, loss.detach(), 1e-5, 1e-6, tag + " loss")
        assert torch.equal(po, pred)
        assert set(go.keys()) == set(grads.keys()), (sorted(go.keys()), sorted(grads.keys()))
        worst = max(float((go[n] - grads[n]).abs().max() / (grads[n].abs().max() + 1e-12)) for n in grads)
        print(f"   worst rel grad err over all tensors: {worst:.3e}")
        assert worst < 1e-3

    def task_boundary(task, x_all):
        loader = [{"image": x_all[i:i + 40]} for i in range(0, x_all.shape[0], 40)]
        torch.manual_seed(900 + task)
        ref.after_task(task, None, loader, None)
        torch.manual_seed(900 + task)
        sel = torch.randperm(x_all.size(0))[:125]
        orc.after_task(x_all[sel])
        ranks = [f.shape[1] for f in ref.feature_list]
        assert ranks == [f.shape[1] for f in orc.feature_list], (ranks, [f.shape[1] for f in orc.feature_list])
        out[f"t{task}/rank"] = np.array(ranks)
        prng = np.random.default_rng(77 + task)
        for i, (fr, fo) in enumerate(zip(ref.feature_list, orc.feature_list)):
            v = prng.standard_normal(fr.shape[0])
            pr, po = fr @ (fr.T @ v), fo @ (fo.T @ v)
            close(po, pr, 1e-5, 1e-6, f"t{task} projector of layer {i} on a probe vector (rank {fr.shape[1]})")
            out[f"t{task}/proj_probe/{i}"] = pr
        print("   ranks", ranks)

    ref.before_task(0, None, None, None); orc.before_task(0)
    x, y = pool[:B], torch.from_numpy(rng.integers(0, 10, (B,)).astype(np.int64))
    both(x, y, "t0s0")
    task_boundary(0, pool[:150])
    ref.before_task(1, None, None, None); orc.before_task(1)
    for s_ in range(2):
        x, y = pool[20 + 16 * s_:36 + 16 * s_], torch.from_numpy(rng.integers(10, 20, (B,)).astype(np.int64))
        both(x, y, f"t1s{s_}")
    task_boundary(1, pool[10:160])
    np.savez_compressed(os.path.join(OUT, "gpm_alexnet.npz"), **out)


def main():
    torch.set_num_threads(8)
    os.makedirs(OUT, exist_ok=True)
    core = import_reference()
    only = [a for a in sys.argv[1:] if not a.startswith("-")]
    if only:                                   # python oracle/make_golden.py inflora [l2p ...]
        for name in only:
            globals()["golden_" + name](core)
        return
    golden_ewc(core)
    golden_icarl(core)
    golden_lwf(core)
    golden_lwf18(core)
    golden_lucir(core)
    golden_herding(core)
    golden_ops(core)
    golden_l2p(core)
    golden_inflora(core)
    golden_dualprompt(core)
    golden_codaprompt(core)
    golden_sdlora(core)
    golden_dualgpm(core)
    golden_inflora_orig(core)
    golden_gpm(core)
    print("golden vectors written to", OUT)


if __name__ == "__main__":
    main()
