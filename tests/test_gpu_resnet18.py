"""GPU: ResNet18 (tiny-imagenet stem, 64 x 64) + LwF — BASELINE config C5's backbone / method pair (`core/model/backbone/resnet.py:26-64,110-246`,
`core/model/lwf.py:28-78`) — through the plugin classes (which call the C ABI), against the CPU oracle on the same seeded inputs and against the golden
vectors written by the REAL `LWF` on the real `resnet18` (tests/golden/lwf_resnet18.npz).

Arithmetic: BF16 GEMM operands (tcgen05 kind::f16), fp32 accumulation, fp32 BatchNorm statistics / residual stream / loss.  Tolerances (fixed):
  vs the oracle in the SAME arithmetic class (conv_mode='bf16'): loss 2e-3, features / logits rel-L2 1e-2 (what differs: summation order, and
      BF16 rounding points that an fp32-ulp difference pushes over — the same discontinuity the TF32 ResNet32 path shows)
  vs the fp32 oracle / the reference golden: loss |d| <= 2e-2, logits rel-L2 <= 3e-2 (SURVEY 8c, bf16 class), per-tensor gradient norm within 25 %,
      whole-arena gradient rel-L2 <= 3.5e-1 (BN + ReLU gradients at random init are discontinuous in the activations: tests/test_gpu_parity_bench_config.py)
  integer outputs (pred): >= 3 of 4 identical (near-tied logits of a random head)."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.golden_util import load, synth_batch, synth_resnet18_state

pytestmark = pytest.mark.gpu
B = 4


def rel_l2(a, b):
    a, b = a.detach().double().cpu().reshape(-1), b.detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_backbone(p, b, max_batch=B):
    import libcontinual_b200.model as M
    bb = M.resnet18(args={"dataset": "tiny-imagenet", "init_cls_num": 20, "inc_cls_num": 10}, max_batch=max_batch, num_classes=200)
    sd = bb.state_dict()
    bb.load_state_dict({**p, **b, "fc.weight": sd["fc.weight"], "fc.bias": sd["fc.bias"]}, strict=True)
    return bb


def grads_of(m):
    eng = m.engine
    d = {"backbone." + n: eng.param_view(n, eng.grads).clone().cpu() for n, _ in eng.layout}
    gw, gb = eng.fc_views(eng.ncls, eng.grads)
    d["classifier.weight"], d["classifier.bias"] = gw.clone().cpu(), gb.clone().cpu()
    return d


def test_resnet18_reference_names_and_shapes():
    p, b, _, _ = synth_resnet18_state(1818, 20)
    bb = make_backbone(p, b)
    names = [n for n, _ in bb.named_parameters()]
    assert names[:3] == ["conv1.0.weight", "conv1.1.weight", "conv1.1.bias"] and names[-2:] == ["fc.weight", "fc.bias"]
    assert "layer2.0.downsample.0.weight" in names and "layer1.0.downsample.0.weight" not in names
    assert sum(q.numel() for n, q in bb.named_parameters() if not n.startswith("fc.")) == 11_168_832      # conv + BN parameters of ResNet18 with a 3x3 stem
    sd = bb.state_dict()
    assert "layer4.1.bn2.running_var" in sd and "conv1.1.num_batches_tracked" in sd
    assert torch.equal(sd["layer3.0.conv1.weight"].cpu(), p["layer3.0.conv1.weight"])


@pytest.mark.parametrize("train", [True, False])
def test_resnet18_forward_vs_oracle(train):
    p, b, _, _ = synth_resnet18_state(1818, 20)
    bb = make_backbone(p, b, max_batch=8)
    x, _ = synth_batch(77, 8, 0, 10, img=64)
    bb.train(train)
    with torch.no_grad():
        out = bb(x.cuda())
    assert not bb.engine.tensor_core_error()
    ob = {k: v.clone() for k, v in b.items()}
    ref_bf = port.resnet18_forward(p, {k: v.clone() for k, v in b.items()}, x, train, True, conv_mode="bf16")
    ref_32 = port.resnet18_forward(p, ob, x, train, True)
    e_bf, e_32 = rel_l2(out["features"], ref_bf["features"]), rel_l2(out["features"], ref_32["features"])
    print(f"resnet18 forward (train={train}): features rel-L2 vs bf16 oracle {e_bf:.2e}, vs fp32 oracle {e_32:.2e}")
    assert e_bf <= 1e-2 and e_32 <= 3e-2
    for a, r in zip(out["fmaps"], ref_32["fmaps"]):
        assert tuple(a.shape) == tuple(r.shape) and rel_l2(a, r) <= 3e-2
    if train:                                  # running statistics moved like nn.BatchNorm2d's
        sd = bb.state_dict()
        for k, v in ob.items():
            if "num_batches" in k:
                assert int(sd[k]) == int(v)
            else:
                assert torch.allclose(sd[k].cpu(), v, rtol=2e-2, atol=2e-3), k


def test_resnet18_backward_composition_with_pinned_relu_masks():
    """The gradient comparison without the mask-flip discontinuity: the oracle (BF16 arithmetic class) is run with every ReLU replaced by a multiplication
    with the mask the CUDA path actually used (read back from its stored activations).  With the routing pinned, the backward is a fixed linear map and
    every parameter gradient must agree to BF16 round-off: per tensor rel-L2 <= 3e-2 (measured: see the printed line), whole arena <= 2e-2.  This is
    the test that would expose a wrong buffer, a missing shortcut gradient or a wrong col2im / split-K reduction, which the 20 % chaos band of the
    free-running comparison would hide."""
    import torch.nn.functional as F
    import libcontinual_b200.model as M
    Bq = 8
    p, b, fc_w, fc_b = synth_resnet18_state(1818, 20)
    bb = make_backbone(p, b, max_batch=Bq)
    m = M.LWF(bb, 512, 200, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10)
    m.before_task(0, None, None, None)
    eng = m.engine
    w, bias = eng.fc_views(10)
    w.copy_(fc_w[:10].cuda()); bias.copy_(fc_b[:10].cuda())
    m.train()
    x, y = synth_batch(1900, Bq, 0, 10, img=64)
    pred, acc, loss = m.observe({"image": x, "label": y})
    torch.cuda.synchronize()
    got = grads_of(m)
    ws = eng.ws
    nchw = lambda t, h, c: t[:Bq * h * h].view(Bq, h, h, c).permute(0, 3, 1, 2).float().cpu()
    masks = {"conv1": nchw(ws.a0, 64, 64) > 0}
    for blk in eng.blocks:
        c2 = blk["conv2"]
        masks[blk["name"] + ".relu1"] = nchw(ws.a1[blk["name"]], c2.Ho, c2.cout) > 0
        masks[blk["name"] + ".relu2"] = nchw(ws.out_f32[blk["name"]], c2.Ho, c2.cout) > 0
    pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    ow, ob_ = fc_w[:10].clone().requires_grad_(True), fc_b[:10].clone().requires_grad_(True)
    feat = port.resnet18_forward(pr, {k: v.clone() for k, v in b.items()}, x, True, True, conv_mode="bf16", relu=lambda name, t: t * masks[name])["features"]
    lo = F.cross_entropy(F.linear(feat, ow, ob_), y)
    gs = torch.autograd.grad(lo, list(pr.values()) + [ow, ob_])
    go = {"backbone." + k: g for k, g in zip(pr.keys(), gs)}
    go["classifier.weight"], go["classifier.bias"] = gs[-2], gs[-1]
    errs = {k: rel_l2(got[k], go[k]) for k in go}
    a = torch.cat([got[k].reshape(-1).double() for k in go]); r = torch.cat([go[k].reshape(-1).double() for k in go])
    whole = float((a - r).norm() / r.norm())
    worst = max(errs, key=errs.get)
    print(f"pinned-mask gradient check: loss {float(loss):.5f} vs {float(lo):.5f}; whole arena rel-L2 {whole:.2e}, median {float(np.median(list(errs.values()))):.2e}, "
          f"worst {worst} {errs[worst]:.2e}")
    assert abs(float(loss) - float(lo)) <= 2e-3
    assert whole <= 2e-2 and errs[worst] <= 3e-2, (whole, worst, errs[worst])


def _check_step(m, orc, orc_bf, x, y, g, tag, opt):
    pred, acc, loss = m.observe({"image": x, "label": y})
    opt.zero_grad(); loss.backward()
    torch.cuda.synchronize()
    assert not m.engine.tensor_core_error()
    got = grads_of(m)
    po, ao, lo, go = orc.step(x, y, apply_update=False)
    pb, ab, lb, gb = orc_bf.step(x, y, apply_update=False)
    flat = lambda d, ref: torch.cat([d[k].reshape(-1).double() for k in ref])
    whole_32 = float((flat(got, go) - flat(go, go)).norm() / flat(go, go).norm())
    whole_bf = float((flat(got, gb) - flat(gb, gb)).norm() / flat(gb, gb).norm())
    print(f"[{tag}] loss {float(loss):.5f} (fp32 oracle {float(lo):.5f}, bf16 oracle {float(lb):.5f}); gradient rel-L2 whole arena vs fp32 {whole_32:.3f}, vs bf16 {whole_bf:.3f}")
    assert abs(float(loss) - float(lb)) <= 2e-3 and abs(float(loss) - float(lo)) <= 2e-2
    assert whole_32 <= 3.5e-1                      # measured 0.29 at B = 4 (chaos band; the pinned-mask test above carries the tight bound)
    assert float((pred.cpu() == po).float().mean()) >= 0.75
    # the reference's own numbers
    assert abs(float(loss) - float(g[tag + "/loss"])) <= 2e-2
    names = [str(n) for n in g[tag + "/grad/names"]]
    bad = []
    for i, n in enumerate(names):
        ref_norm = float(g[tag + "/grad/norm"][i])
        if abs(float(got[n].double().norm()) - ref_norm) > 0.25 * ref_norm + 1e-6:
            bad.append(n)
    assert len(bad) <= len(names) // 20, bad


def test_lwf_resnet18_steps_vs_oracle_and_reference_golden():
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    g = load("lwf_resnet18.npz")
    p, b, fc_w, fc_b = synth_resnet18_state(1818, 20)
    bb = make_backbone(p, b)
    m = M.LWF(bb, 512, 200, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10)
    m.before_task(0, None, None, None)
    eng = m.engine
    w, bias = eng.fc_views(10)
    w.copy_(fc_w[:10].cuda()); bias.copy_(fc_b[:10].cuda())
    m.train()
    opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=eng)
    mk = lambda mode: port.ResNetMethodOracle("lwf", p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10, arch="resnet18", maxpool=True, conv_mode=mode)
    orc, orc_bf = mk("fp32"), mk("bf16")
    x, y = synth_batch(1900, B, 0, 10, img=64)
    _check_step(m, orc, orc_bf, x, y, g, "t0s0", opt)
    opt.step()

    def sync(o):
        with torch.no_grad():
            for n in o.p:
                o.p[n].copy_(eng.param_view(n).cpu())
            ww, bb_ = eng.fc_views(eng.ncls)
            o.fc_w.copy_(ww.cpu()); o.fc_b.copy_(bb_.cpu())
            for bn in eng.bn_names:
                mu, var = eng.running_views(bn)
                o.b[bn + ".running_mean"].copy_(mu.cpu()); o.b[bn + ".running_var"].copy_(var.cpu())
    sync(orc); sync(orc_bf)
    m.before_task(1, None, None, None)          # update_fc + frozen copy of backbone and head (lwf.py:28-50)
    w, bias = eng.fc_views(20)
    w[10:].copy_(fc_w[10:20].cuda()); bias[10:].copy_(fc_b[10:20].cuda())
    for o in (orc, orc_bf):
        o.snapshot_teacher(); o.prev_cls = 10; o.task_idx = 1
        o.grow_head(fc_w[:20], fc_b[:20]); o.reset_optimizer()
    opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=eng)
    gold_free = {k: g[k] for k in g.files}
    for s in range(2):
        x, y = synth_batch(1910 + s, B, 10, 20, img=64)
        # from step t1s0 on the golden trajectory has diverged by the first step's rounding: compare against the oracles from synced state, and against
        # the golden's loss only loosely (same data, nearly the same weights)
        pred, acc, loss = m.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward()
        torch.cuda.synchronize()
        assert float(eng.scal[3]) > 0.0                                   # KD term live
        got = grads_of(m)
        _, _, lo, go = orc.step(x, y, apply_update=False)
        _, _, lb, gb = orc_bf.step(x, y, apply_update=False)
        flat = lambda d, ref: torch.cat([d[k].reshape(-1).double() for k in ref])
        whole_32 = float((flat(got, go) - flat(go, go)).norm() / flat(go, go).norm())
        print(f"[t1s{s}] loss {float(loss):.5f} (fp32 oracle {float(lo):.5f}, bf16 oracle {float(lb):.5f}, reference {float(gold_free[f't1s{s}/loss']):.5f}); "
              f"gradient rel-L2 whole arena vs fp32 {whole_32:.3f}")
        assert abs(float(loss) - float(lb)) <= 5e-3 and abs(float(loss) - float(lo)) <= 3e-2 and whole_32 <= 3.5e-1
        opt.step()
        sync(orc); sync(orc_bf)


def test_lwf_resnet18_graphed_step_equals_eager():
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    from libcontinual_b200.trainer import GraphedStep, train_step_eager
    outs = []
    for graphed in (False, True):
        torch.manual_seed(0)                 # the teacher snapshot of before_task(1) contains head rows drawn from the torch RNG
        p, b, fc_w, fc_b = synth_resnet18_state(99, 120)
        bb = make_backbone(p, b, max_batch=8)
        m = M.LWF(bb, 512, 200, device=torch.device("cuda"), init_cls_num=100, inc_cls_num=20)
        m.before_task(0, None, None, None)
        m.before_task(1, None, None, None)
        w, bias = m.engine.fc_views(120)
        w.copy_(fc_w.cuda()); bias.copy_(fc_b.cuda())
        m.train()
        opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=m.engine)
        step = GraphedStep(m, opt, 8) if graphed else None
        for s in range(2):
            x, y = synth_batch(600 + s, 8, 100, 120, img=64)
            if graphed:
                step.run(x.cuda(), y.cuda())
            else:
                train_step_eager(m, opt, {"image": x, "label": y})
        torch.cuda.synchronize()
        outs.append((m.engine.params.clone(), m.engine.rstat.clone(), float(m.engine.scal[0])))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


def test_lwf_resnet18_with_gradient_projection_c5():
    """BASELINE config C5 ("LwF + GPM" on ResNet18): the LwF task-1 step followed by the GPM projection of every 3x3 / 1x1 conv gradient onto the
    complement of seeded orthonormal bases (rank = 10 % of Cin*k*k).  Checks: (1) the projected gradient equals the SAME step's unprojected
    gradient projected in float64 (the projection operator alone, 2e-5); (2) it is orthogonal to the bases; (3) non-projected tensors are untouched;
    (4) the step through the oracle with `proj` agrees to the LwF tolerance."""
    import libcontinual_b200.model as M
    torch.manual_seed(3)
    p, b, fc_w, fc_b = synth_resnet18_state(2020, 20)

    def build():
        torch.manual_seed(3)
        bb = make_backbone(p, b)
        m = M.LWF(bb, 512, 200, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10)
        m.before_task(0, None, None, None)
        w, bias = m.engine.fc_views(10)
        w.copy_(fc_w[:10].cuda()); bias.copy_(fc_b[:10].cuda())
        m.before_task(1, None, None, None)
        w, bias = m.engine.fc_views(20)
        w[10:].copy_(fc_w[10:20].cuda()); bias[10:].copy_(fc_b[10:20].cuda())
        m.train()
        return m

    m0, m1 = build(), build()
    eng = m1.engine
    rng = np.random.default_rng(55)
    bases = {}
    for name, shape in eng.layout:
        if len(shape) == 4 and (shape[1] * shape[2] * shape[3]) % 8 == 0:
            D = shape[1] * shape[2] * shape[3]
            q, _ = np.linalg.qr(rng.standard_normal((D, max(1, D // 10))))
            bases[name] = torch.from_numpy(q.astype(np.float32))
    assert len(bases) >= 19                                   # every conv but the 3-channel stem
    m1.set_gradient_projection(bases)
    x, y = synth_batch(2100, B, 10, 20, img=64)
    m0.observe({"image": x, "label": y})
    m1.observe({"image": x, "label": y})
    torch.cuda.synchronize()
    assert not eng.tensor_core_error()
    for name, shape in eng.layout:
        g0 = m0.engine.param_view(name, m0.engine.grads).double().cpu()
        g1 = eng.param_view(name, eng.grads).double().cpu()
        if name in bases:
            U = bases[name].double()
            want = g0.view(shape[0], -1) - (g0.view(shape[0], -1) @ U) @ U.T
            assert rel_l2(g1.view(shape[0], -1), want) < 2e-5, name
            assert float((g1.view(shape[0], -1) @ U).norm() / (g0.view(shape[0], -1) @ U).norm()) < 1e-4, name
        else:
            assert torch.equal(g0, g1), name
    orc = port.ResNetMethodOracle("lwf", p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10, arch="resnet18", maxpool=True, conv_mode="bf16")
    orc.snapshot_teacher(); orc.prev_cls = 10; orc.task_idx = 1
    orc.grow_head(fc_w[:20], fc_b[:20])
    orc.proj = {"backbone." + n: (U @ U.T) for n, U in bases.items()}
    _, _, lo, go = orc.step(x, y, apply_update=False)
    assert abs(float(eng.scal[0]) - float(lo)) <= 5e-3
    got = torch.cat([eng.param_view(n, eng.grads).reshape(-1).double().cpu() for n in bases])
    ref = torch.cat([go["backbone." + n].reshape(-1).double() for n in bases])
    assert float((got - ref).norm() / ref.norm()) <= 3.5e-1
