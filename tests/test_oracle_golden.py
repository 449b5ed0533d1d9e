"""CPU: the oracle restatement (oracle/port.py) reproduces the golden vectors that oracle/make_golden.py recorded
from the real reference classes (imported from /root/reference in the build container)."""
import numpy as np
import torch
import torch.nn.functional as F

from oracle import port
from tests.golden_util import check_summary, load, synth_batch, synth_resnet_state

B = 8


def _run(orc, g, tag, x, y, rtol=1e-5):
    pred, acc, loss, grads = orc.step(x, y)
    assert abs(float(loss) - float(g[tag + "/loss"])) <= 1e-5 * abs(float(g[tag + "/loss"])) + 1e-6
    assert np.array_equal(pred.numpy(), g[tag + "/pred"])
    check_summary(g, tag + "/grad", grads, rtol=rtol)


def test_ewc_trajectory_matches_reference():
    g = load("ewc_resnet32.npz")
    p, b, fc_w, fc_b = synth_resnet_state(101, 20)
    orc = port.ResNetMethodOracle("ewc", p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10, lamda=1000.0)
    for s in range(2):
        _run(orc, g, f"t0s{s}", *synth_batch(1000 + s, B, 0, 10))
    check_summary(g, "t0/param", orc.named())
    fb = [synth_batch(1100 + i, B if i < 2 else 5, 0, 10) for i in range(3)]
    orc.ewc_after_task(fb, B)
    check_summary(g, "t0/fisher", orc.fisher, rtol=2e-4, atol=1e-10)
    orc.task_idx = 1
    orc.grow_head(fc_w[:20], fc_b[:20]); orc.reset_optimizer()
    for s in range(3):
        _run(orc, g, f"t1s{s}", *synth_batch(1200 + s, B, 10, 20))
    check_summary(g, "t1/param", orc.named())
    check_summary(g, "t1/bnbuf", {"backbone." + k: v.float() for k, v in orc.b.items()})
    orc.ewc_after_task([synth_batch(1300 + i, B, 10, 20) for i in range(2)], B)
    check_summary(g, "t1/fisher", orc.fisher, rtol=2e-4, atol=1e-10)


def test_ewc_penalty_is_active_in_task1():
    g = load("ewc_resnet32.npz")
    # first task-1 step has theta == theta*, so loss is pure CE; the next ones carry lamda * penalty
    assert float(g["t1s1/loss"]) > 0 and float(g["t1s2/loss"]) > 0


def test_icarl_trajectory_matches_reference():
    g = load("icarl_resnet32.npz")
    p, b, fc_w, fc_b = synth_resnet_state(202, 100)
    orc = port.ResNetMethodOracle("icarl", p, b, fc_w, fc_b, init_cls=10, inc_cls=5)
    for s in range(2):
        _run(orc, g, f"t0s{s}", *synth_batch(2000 + s, B, 0, 10))
    orc.snapshot_teacher(); orc.prev_cls = 10; orc.accu_cls = 15; orc.task_idx = 1; orc.reset_optimizer()
    for s in range(2):
        _run(orc, g, f"t1s{s}", *synth_batch(2100 + s, B, 0, 15))
    check_summary(g, "t1/param", {"backbone." + k: v for k, v in orc.p.items()} | {"classifier.weight": orc.fc_w, "classifier.bias": orc.fc_b})


def test_lwf_trajectory_matches_reference():
    g = load("lwf_resnet32.npz")
    p, b, fc_w, fc_b = synth_resnet_state(303, 20)
    orc = port.ResNetMethodOracle("lwf", p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10)
    _run(orc, g, "t0s0", *synth_batch(3000, B, 0, 10))
    orc.snapshot_teacher(); orc.prev_cls = 10; orc.task_idx = 1
    orc.grow_head(fc_w[:20], fc_b[:20]); orc.reset_optimizer()
    for s in range(2):
        _run(orc, g, f"t1s{s}", *synth_batch(3100 + s, B, 10, 20))


def test_lwf_resnet18_trajectory_matches_reference():
    """BASELINE config C5's pair: the oracle's resnet18 (tiny-imagenet stem, 64 x 64) + LwF against the golden written by the real `LWF` on the real
    `resnet18` (oracle/make_golden.py::golden_lwf18)."""
    from tests.golden_util import synth_resnet18_state
    g = load("lwf_resnet18.npz")
    p, b, fc_w, fc_b = synth_resnet18_state(1818, 20)
    orc = port.ResNetMethodOracle("lwf", p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10, arch="resnet18", maxpool=True)
    x, y = synth_batch(1900, 4, 0, 10, img=64)
    _run(orc, g, "t0s0", x, y)
    with torch.no_grad():
        f = port.resnet18_forward({k: v.detach() for k, v in orc.p.items()}, orc.b, x, True, True)["features"]
    assert np.allclose(f.numpy(), g["t0s0/features_after"], rtol=1e-4, atol=1e-5)
    orc.snapshot_teacher(); orc.prev_cls = 10; orc.task_idx = 1
    orc.grow_head(fc_w[:20], fc_b[:20]); orc.reset_optimizer()
    for s in range(2):
        _run(orc, g, f"t1s{s}", *synth_batch(1910 + s, 4, 10, 20, img=64))


def test_gpm_alexnet_matches_reference():
    """`oracle.port.GPMOracle` (AlexNet_TRGP forward, projected gradients, basis construction / growth) against the golden written by the real `GPM`
    on the real `AlexNet_TRGP` (oracle/make_golden.py::golden_gpm): losses, predictions, per-tensor gradient norms, basis ranks, projectors on probe vectors."""
    from tests.golden_util import synth_alexnet_state
    g = load("gpm_alexnet.npz")
    p, heads = synth_alexnet_state(4040)
    orc = port.GPMOracle(p, heads, 10, 10, lr=0.01)
    rng = np.random.default_rng(4141)
    pool = torch.from_numpy(rng.standard_normal((160, 3, 32, 32)).astype(np.float32))

    def step(x, y, tag):
        pred, acc, loss, grads = orc.step(x, y)
        assert abs(float(loss) - float(g[tag + "/loss"])) <= 1e-5 * abs(float(g[tag + "/loss"])) + 1e-6
        assert np.array_equal(pred.numpy(), g[tag + "/pred"])
        names = [str(n) for n in g[tag + "/grad/names"]]
        assert sorted(names) == sorted(grads.keys())
        for i, n in enumerate(names):
            ref = float(g[tag + "/grad/norm"][i])
            assert abs(float(grads[n].double().norm()) - ref) <= 1e-4 * ref + 1e-7, (tag, n)

    def boundary(task, x_all):
        torch.manual_seed(900 + task)
        sel = torch.randperm(x_all.size(0))[:125]
        orc.after_task(x_all[sel])
        assert [f.shape[1] for f in orc.feature_list] == list(g[f"t{task}/rank"])
        prng = np.random.default_rng(77 + task)
        for i, f in enumerate(orc.feature_list):
            v = prng.standard_normal(f.shape[0])
            assert np.allclose(f @ (f.T @ v), g[f"t{task}/proj_probe/{i}"], rtol=1e-4, atol=1e-5)

    orc.before_task(0)
    step(pool[:16], torch.from_numpy(rng.integers(0, 10, (16,)).astype(np.int64)), "t0s0")
    boundary(0, pool[:150])
    orc.before_task(1)
    for s in range(2):
        step(pool[20 + 16 * s:36 + 16 * s], torch.from_numpy(rng.integers(10, 20, (16,)).astype(np.int64)), f"t1s{s}")
    boundary(1, pool[10:160])


def test_l2p_select_matches_reference():
    g = load("ops_small.npz")
    rng = np.random.default_rng(404)
    for case in range(4):
        Bq, pool, topk, length, D = [int(v) for v in g[f"l2p{case}/shape"]]
        prm = torch.from_numpy(rng.uniform(0, 1, (1, pool, length, D)).astype(np.float32))
        key = torch.from_numpy(rng.uniform(0, 1, (pool, D)).astype(np.float32))
        q = torch.from_numpy(rng.standard_normal((Bq, D)).astype(np.float32))
        if case == 3:
            key[1] = key[0]; key[4] = key[0]
        key.requires_grad_(True)
        bp, rs, ids = port.l2p_select(prm, key, q, topk)
        assert np.array_equal(ids.numpy(), g[f"l2p{case}/ids"])
        assert abs(float(rs) - float(g[f"l2p{case}/reduce_sim"])) < 1e-6
        assert abs(float(bp.double().sum()) - float(g[f"l2p{case}/prompt_sum"])) < 1e-6 * abs(float(g[f"l2p{case}/prompt_sum"]))
        (-rs).backward()
        assert np.allclose(key.grad.numpy(), g[f"l2p{case}/dkey"], rtol=1e-5, atol=1e-7)
        if f"l2p{case}/strict" in g.files and int(g[f"l2p{case}/strict"]):
            kn = F.normalize(key.detach(), dim=-1); qn = F.normalize(q, dim=-1)
            assert np.array_equal(port.l2p_majority_ids_numpy((qn @ kn.T).numpy(), topk), g[f"l2p{case}/ids"])
    # the draws for the cosine-head fixtures follow in the same stream
    feat = torch.from_numpy(rng.standard_normal((16, 64)).astype(np.float32))
    w1 = torch.from_numpy(rng.uniform(-0.125, 0.125, (10, 64)).astype(np.float32))
    w2 = torch.from_numpy(rng.uniform(-0.125, 0.125, (5, 64)).astype(np.float32))
    assert np.allclose(port.cosine_head(feat, w1, torch.tensor([1.7])).numpy(), g["cos/out"], rtol=1e-6, atol=1e-7)
    assert np.allclose(port.cosine_head(feat, torch.cat([w1, w2]), torch.tensor([2.5])).numpy(), g["cos/split_out"], rtol=1e-6, atol=1e-7)


def test_known_answer_properties():
    """Reference-free properties listed in SURVEY.md §4."""
    rng = np.random.default_rng(7)
    t = torch.from_numpy(rng.standard_normal((6, 9)).astype(np.float32))
    s = t.clone().requires_grad_(True)
    port.kd_loss(s, t, 2.0).backward()
    assert float(s.grad.abs().max()) < 1e-7                       # KD minimised at s == t
    U = torch.linalg.qr(torch.from_numpy(rng.standard_normal((32, 5)).astype(np.float32)))[0]
    gproj = port.gpm_project(torch.from_numpy(rng.standard_normal((7, 32)).astype(np.float32)), U @ U.T)
    assert float((gproj @ U).abs().max()) < 1e-5                  # projected gradient is orthogonal to the stored basis
    feat = torch.from_numpy(rng.standard_normal((4, 64)).astype(np.float32))
    w = torch.from_numpy(rng.standard_normal((10, 64)).astype(np.float32))
    assert float(port.cosine_head(feat, w, torch.tensor([3.0])).abs().max()) <= 3.0 + 1e-5
    qkv = torch.from_numpy(rng.standard_normal((12, 4)).astype(np.float32))
    A = torch.from_numpy(rng.standard_normal((2, 4)).astype(np.float32))
    assert torch.equal(port.lora_merge_qkv(qkv, A, torch.zeros(4, 2), A, torch.zeros(4, 2)), qkv)   # B = 0 -> frozen model


def lucir_oracle_step(p, b, W, sigma, x, y, teacher, n_old, cur_lamda, K=2, dist=0.5, lw_mr=1.0):
    """One LUCIR loss evaluation with the oracle pieces (lucir.py:175-205).  `teacher` = (params, buffers) of the frozen copy or
    None (task 0: CE only).  Returns (loss, pred, grads dict keyed like the product's arena: backbone.<n>, head, sigma)."""
    op = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    Wv, sv = W.clone().requires_grad_(True), sigma.clone().requires_grad_(True)
    rfeat = None
    if teacher is not None:
        with torch.no_grad():
            rfeat = port.cifar_resnet_forward(teacher[0], {k: v.clone() for k, v in teacher[1].items()}, x, False, last_relu=False)["features"]
    feat = port.cifar_resnet_forward(op, b, x, True, last_relu=False)["features"]
    scores = port.cosine_head(feat, Wv, None)
    logits = sv * scores
    loss = F.cross_entropy(logits, y) if teacher is None else port.lucir_loss(feat, rfeat, logits, scores, y, n_old, cur_lamda, K, dist, lw_mr)
    gs = torch.autograd.grad(loss, list(op.values()) + [Wv, sv])
    grads = {"backbone." + k: g for k, g in zip(op.keys(), gs)}
    grads["head"], grads["sigma"] = gs[-2], gs[-1]
    return loss.detach(), logits.argmax(1), grads


def test_lucir_matches_reference():
    from tests.golden_util import synth_resnet_state
    g = load("lucir_resnet32.npz")
    p, b, fc_w, fc_b = synth_resnet_state(404, 15)
    bb = {k: v.clone() for k, v in b.items()}
    x, y = synth_batch(4100, 8, 0, 10)
    loss, pred, grads = lucir_oracle_step(p, bb, fc_w[:10], torch.tensor([1.5]), x, y, None, 0, 5.0)
    assert abs(float(loss) - float(g["t0/loss"])) < 1e-6 and np.array_equal(pred.numpy(), g["t0/pred"])
    teacher = (p, {k: v.clone() for k, v in bb.items()})
    x, y = synth_batch(4101, 8, 0, 15)
    y[0], y[1] = 3, 12
    assert np.array_equal(y.numpy(), g["t1/y"])
    loss, pred, grads = lucir_oracle_step(p, bb, fc_w[:15], torch.tensor([1.5]), x, y, teacher, 10, float(g["t1/cur_lamda"]))
    assert abs(float(loss) - float(g["t1/loss"])) < 1e-5 and np.array_equal(pred.numpy(), g["t1/pred"])
    assert np.allclose(grads["head"][:10].numpy(), g["t1/full/classifier.fc1.weight"], rtol=1e-4, atol=1e-7)
    assert np.allclose(grads["head"][10:].numpy(), g["t1/full/classifier.fc2.weight"], rtol=1e-4, atol=1e-7)
    assert np.allclose(grads["sigma"].numpy(), g["t1/full/classifier.sigma"], rtol=1e-4, atol=1e-7)


def test_herding_matches_reference():
    from oracle.make_golden import herding_inputs
    raw, labels = herding_inputs()
    feats = raw / raw.norm(dim=1).view(-1, 1)
    assert port.herding_select(feats, labels, 25) == [int(i) for i in load("herding.npz")["idx"]]


def test_l2p_vit_observe_matches_reference():
    """Oracle ViT-B/16 + L2P observe() vs the real `core.model.l2p.L2P` on `vit_pt_imnet` (fixture: tests/golden/l2p_vit.npz).
    One task only here (the generator checked both) to keep the CPU suite short."""
    from tests.golden_util import l2p_oracle_step, synth_images, synth_vit_state
    g = load("l2p_vit.npz")
    torch.set_num_threads(8)
    p, prm, key, fc_w, fc_b = synth_vit_state(5150)
    x, y = synth_images(601, 4, 10, 20)
    o = l2p_oracle_step(p, prm, key, fc_w, fc_b, x, y, 10, 20)
    assert np.array_equal(o["major"].numpy(), g["t1/major"])
    assert abs(float(o["loss"]) - float(g["t1/loss"])) < 1e-5
    for k in ("logits", "dprompt", "dkey", "dW", "db", "cls_features", "feat"):
        ref = torch.from_numpy(g["t1/" + k])
        err = float((o[k] - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < 1e-4, (k, err)


def test_inflora_vit_observe_matches_reference():
    """Oracle InfLoRA_OPT step (weight-side adapters on k, v; CE over the task head) vs the real `core.model.InfLoRA_opt.InfLoRA_OPT`
    (fixture: tests/golden/inflora_vit.npz), task 0 only here (the generator checked task 1 on the merged weights as well)."""
    import torch.nn.functional as F
    from tests.golden_util import synth_images, synth_lora_state, synth_vit_state
    g = load("inflora_vit.npz")
    torch.set_num_threads(8)
    p = synth_vit_state(5150)[0]
    lora, hw, hb = synth_lora_state(880)
    x, y = synth_images(700, 4, 0, 20)
    ol = [{k: v.clone().requires_grad_(k.startswith("B_")) for k, v in d.items()} for d in lora]
    ow = hw.clone().requires_grad_(True); ob = hb.clone().requires_grad_(True)
    logits = port.inflora_logits(p, ol, ow, ob, x)
    loss = F.cross_entropy(logits, y)
    loss.backward()
    assert abs(float(loss) - float(g["t0/loss"])) < 1e-5
    got = {"logits": logits.detach(), "dW": ow.grad, "db": ob.grad, "dB_k": torch.stack([d["B_k"].grad for d in ol]),
           "dB_v": torch.stack([d["B_v"].grad for d in ol])}
    for k, v in got.items():
        ref = torch.from_numpy(g["t0/" + k])
        err = float((v - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < 1e-4, (k, err)


def test_dualprompt_vit_observe_matches_reference():
    """Oracle DualPrompt step (prefix keys / values on blocks 0-4, task-id bootstrap key loss, masked CE) vs the real reference pool + ViT blocks +
    classifier (fixture: tests/golden/dualprompt_vit.npz), task 1 only here (the generator checked task 0 and the inference selection as well)."""
    from tests.golden_util import synth_dual_pool, synth_images, synth_vit_state
    g = load("dualprompt_vit.npz")
    torch.set_num_threads(8)
    p = synth_vit_state(5150)[0]
    pool, fc_w, fc_b = synth_dual_pool(930)
    x, y = synth_images(751, 4, 10, 20)
    op = {k: v.clone().requires_grad_(True) for k, v in pool.items()}
    ow = fc_w[:20].clone().requires_grad_(True); ob = fc_b[:20].clone().requires_grad_(True)
    feat, ploss, q, _ = port.dualprompt_forward(p, op, x, 1, True)
    loss = port.dualprompt_loss(port.linear_head(feat, ow, ob), y, 10, ploss)
    loss.backward()
    assert abs(float(loss) - float(g["t1/loss"])) < 1e-4 and abs(float(ploss) - float(g["t1/ploss"])) < 1e-4
    got = {"dW": ow.grad, "db": ob.grad, "feat": feat.detach(), "query": q}
    got.update({"d" + k: v.grad for k, v in op.items()})
    for k, v in got.items():
        ref = torch.from_numpy(g["t1/" + k])
        err = float((v - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < 2e-4, (k, err)


def test_codaprompt_vit_observe_matches_reference():
    """Oracle CodaPrompt step (attention-weighted prompt components as prefix keys / values on blocks 0-4, masked CE) vs the real reference pool + ViT
    blocks + classifier (fixture: tests/golden/codaprompt_vit.npz), task 1 only here."""
    from tests.golden_util import synth_coda_pool, synth_images, synth_vit_state
    g = load("codaprompt_vit.npz")
    torch.set_num_threads(8)
    p = synth_vit_state(5150)[0]
    pool, fc_w, fc_b = synth_coda_pool(940)
    x, y = synth_images(761, 4, 10, 20)
    op = {k: v.clone().requires_grad_(True) for k, v in pool.items()}
    ow = fc_w[:20].clone().requires_grad_(True); ob = fc_b[:20].clone().requires_grad_(True)
    feat, q = port.codaprompt_forward(p, op, x, 10)
    loss = port.dualprompt_loss(port.linear_head(feat, ow, ob), y, 10, torch.zeros(()))
    loss.backward()
    assert abs(float(loss) - float(g["t1/loss"])) < 1e-5
    got = {"dW": ow.grad, "db": ob.grad, "feat": feat.detach()}
    got.update({"d" + k: v.grad[:10] for k, v in op.items()})
    for k, v in got.items():
        ref = torch.from_numpy(g["t1/" + k])
        err = float((v - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < 1e-3, (k, err)


def test_coda_pool_init_matches_reference_draws():
    """Same torch-RNG draws as the reference constructor (uniform + Gram-Schmidt of the first pool/n_tasks rows, zeros elsewhere: prompt.py:48-64,98-144)."""
    from libcontinual_b200.model.codaprompt import CodaPromptPool
    g = load("codaprompt_vit.npz")
    torch.manual_seed(7)
    pool = CodaPromptPool(768, 10, [100, 8, 0.0])
    gram = (pool.e_k_0[:10] @ pool.e_k_0[:10].T).detach().numpy()
    assert np.allclose(gram, np.eye(10), atol=1e-5) and np.allclose(gram, g["init/e_k_0_gram"], atol=1e-5)
    assert float(pool.e_p_3[10:].abs().max()) == float(g["init/e_p_3_tail_absmax"]) == 0.0
    assert np.allclose(pool.e_a_2[:10, :8].detach().numpy(), g["init/e_a_2_head"], atol=1e-6)
    assert np.allclose(pool.e_p_4[:10, 3, :8].detach().numpy(), g["init/e_p_4_head"], atol=1e-6)


def test_sdlora_vit_observe_matches_reference():
    """Oracle SD-LoRA step (scaled sum of per-task q / v adapters, shared trainable magnitudes) vs the real `core.model.sd_lora.SD_LoRA`
    (fixture: tests/golden/sdlora_vit.npz), task 2 (two frozen adapters + the current one)."""
    import torch.nn.functional as F
    from tests.golden_util import sdlora_task_states, synth_images, synth_vit_state
    g = load("sdlora_vit.npz")
    torch.set_num_threads(8)
    p = synth_vit_state(5150)[0]
    blocks, mags, hw, hb = sdlora_task_states(2)[2]
    x, y = synth_images(782, 4, 20, 30)
    ob_ = [[{k: v.clone().requires_grad_(i == 2) for k, v in ad.items()} for i, ad in enumerate(blocks[l])] for l in range(12)]
    om = [mags[i:i + 1].clone().requires_grad_(True) for i in range(3)]
    ow = hw.clone().requires_grad_(True); obias = hb.clone().requires_grad_(True)
    logits = port.sdlora_logits(p, ob_, om, ow, obias, x)
    loss = F.cross_entropy(logits[:, 20:], y - 20)
    loss.backward()
    assert abs(float(loss) - float(g["t2/loss"])) < 1e-5
    got = {"logits": logits.detach(), "dW": ow.grad, "db": obias.grad, "dmag": torch.cat([m.grad for m in om])}
    for nm, key in (("dA_q", "A_q"), ("dB_q", "B_q"), ("dA_v", "A_v"), ("dB_v", "B_v")):
        got[nm] = torch.stack([ob_[l][2][key].grad for l in range(12)])
    for k, v in got.items():
        ref = torch.from_numpy(g["t2/" + k])
        err = float((v - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < 1e-3, (k, err)


def test_inflora_orig_observe_matches_reference():
    """Oracle InfLoRA (original: sum of the per-task k / v adapters, timm-style ViT with LayerNorm eps 1e-6) vs the real `core.model.InfLoRA.InfLoRA`
    on the real `ViT_lora_co` (fixture: tests/golden/inflora_orig_vit.npz), task 1 (two stacked adapters, the second one trains)."""
    import torch.nn.functional as F
    from tests.golden_util import synth_images, synth_stacked_adapters, synth_timm_vit_state
    g = load("inflora_orig_vit.npz")
    torch.set_num_threads(8)
    p, _ = synth_timm_vit_state(5150)
    blocks, hw, hb = synth_stacked_adapters(990, 2)
    x, y = synth_images(801, 4, 10, 20)
    ob_ = [[{k: v.clone().requires_grad_(i == 1 and k.startswith("B_")) for k, v in ad.items()} for i, ad in enumerate(blocks[l])] for l in range(12)]
    ow = hw[1].clone().requires_grad_(True); obias = hb[1].clone().requires_grad_(True)
    logits = port.inflora_orig_logits(p, ob_, ow, obias, x)
    loss = F.cross_entropy(logits, y - 10)
    loss.backward()
    assert abs(float(loss) - float(g["t1/loss"])) < 1e-5
    got = {"logits": logits.detach(), "dW": ow.grad, "db": obias.grad, "dB_k": torch.stack([ob_[l][1]["B_k"].grad for l in range(12)]),
           "dB_v": torch.stack([ob_[l][1]["B_v"].grad for l in range(12)])}
    for k, v in got.items():
        ref = torch.from_numpy(g["t1/" + k])
        err = float((v - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < 1e-3, (k, err)
