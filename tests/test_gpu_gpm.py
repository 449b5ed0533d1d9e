"""GPU: GPM on AlexNet_TRGP (`core/model/gpm.py:43-206`, `core/model/backbone/alexnet.py:94-156`) through the plugin class (which calls the C ABI), against
the CPU oracle on the same seeded inputs and against the golden vectors written by the REAL `GPM` on the real `AlexNet_TRGP` (tests/golden/gpm_alexnet.npz).

Arithmetic: BF16 GEMM operands (tcgen05), fp32 accumulation, fp32 BatchNorm / loss / gradients; the projection g - g (U U^T) at fp32-level accuracy through
a two-term BF16 split.  Tolerances (fixed): loss |d| <= 5e-3 vs the BF16-class oracle and <= 2e-2 vs fp32 oracle / reference; per-tensor gradient rel-L2
<= 6e-2 vs the BF16-class oracle (five layers: little mask-flip chaos), gradient norms within 10 % of the reference's; predictions >= 14 of 16 identical;
projected gradients orthogonal to the stored bases to 1e-4 relative; basis ranks within +-3 of the reference's and projectors within 5e-2 relative on probe
vectors (the representation matrices carry the forward's BF16 rounding, ranks are threshold counts on their spectra)."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.golden_util import load, synth_alexnet_state

pytestmark = pytest.mark.gpu
B = 16


def rel_l2(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu().reshape(-1), torch.as_tensor(b).detach().double().cpu().reshape(-1)
    return float((a - b).norm() / (b.norm() + 1e-30))


def make_model(p, heads, max_batch=160):
    import libcontinual_b200.model as M
    bb = M.AlexNet_TRGP(max_batch=max_batch)
    m = M.GPM(bb, torch.device("cuda"), init_cls_num=10, inc_cls_num=10, task_num=3)
    sd = {"network.backbone." + k: v for k, v in p.items()}
    sd.update({f"network.classifiers.{t}.weight": h for t, h in enumerate(heads)})
    m.load_state_dict(sd, strict=True)
    return m


def grads_of(m):
    eng = m.engine
    d = {n: eng.param_view(n, eng.theta_grad).clone().cpu() for n, _ in eng.layout if not (m._bn_frozen and n.startswith("bn"))}
    d[f"classifiers.{m.cur_task}.weight"] = eng.head_view(m.cur_task, eng.theta_grad).clone().cpu()
    return d


def _pool():
    rng = np.random.default_rng(4141)
    return rng, torch.from_numpy(rng.standard_normal((160, 3, 32, 32)).astype(np.float32))


def _step_check(m, orcs, x, y, g, tag):
    pred, acc, loss = m.observe({"image": x, "label": y})
    torch.cuda.synchronize()
    assert not m.engine.tensor_core_error()
    got = grads_of(m)
    res = {}
    for mode, orc in orcs.items():
        po, ao, lo, go = orc.step(x, y, apply_update=False)
        errs = {k: rel_l2(got[k], go[k]) for k in go}
        res[mode] = (float(lo), errs, po)
        assert set(go.keys()) == set(got.keys())
    lo_bf, e_bf, _ = res["bf16"]
    lo_32, e_32, po = res["fp32"]
    print(f"[{tag}] loss {float(loss):.5f} (bf16 oracle {lo_bf:.5f}, fp32 oracle {lo_32:.5f}, reference {float(g[tag + '/loss']):.5f}); gradient rel-L2 vs bf16 oracle: "
          f"max {max(e_bf.values()):.3e} ({max(e_bf, key=e_bf.get)}), vs fp32 oracle max {max(e_32.values()):.3e}")
    assert abs(float(loss) - lo_bf) <= 5e-3 and abs(float(loss) - lo_32) <= 2e-2 and abs(float(loss) - float(g[tag + "/loss"])) <= 2e-2
    assert max(e_bf.values()) <= 6e-2, e_bf
    assert float((pred.cpu() == po).float().mean()) >= 14 / 16
    names = [str(n) for n in g[tag + "/grad/names"]]
    for i, n in enumerate(names):
        ref = float(g[tag + "/grad/norm"][i])
        assert abs(float(got[n].double().norm()) - ref) <= 0.1 * ref + 1e-7, (tag, n, float(got[n].norm()), ref)
    return got


def test_gpm_two_tasks_vs_oracle_and_reference_golden():
    g = load("gpm_alexnet.npz")
    p, heads = synth_alexnet_state(4040)
    m = make_model(p, heads)
    m.network.eval()                                       # dropout off, like the golden run (BatchNorm uses batch statistics in either mode)
    orcs = {mode: port.GPMOracle(p, heads, 10, 10, lr=0.01, gemm_mode=mode) for mode in ("fp32", "bf16")}
    rng, pool = _pool()
    names = [n for n, _ in m.network.named_parameters()]
    assert names[:3] == ["backbone.conv1.weight", "backbone.bn1.weight", "backbone.bn1.bias"] and names[-1] == "classifiers.2.weight"
    m.before_task(0, None, None, None)
    for o in orcs.values():
        o.before_task(0)
    x, y = pool[:B], torch.from_numpy(rng.integers(0, 10, (B,)).astype(np.int64))
    _step_check(m, orcs, x, y, g, "t0s0")
    # the SGD step of the trainer (trainer.py:606) through the flat optimizer, then the oracles follow the CUDA path's parameters
    from libcontinual_b200.optim import FlatSGD
    opt = FlatSGD(m.get_parameters(None), lr=0.01, model=m)
    before, gsnap = m.theta.clone(), m.theta_grad.clone()
    opt.step()
    lo1, hi1 = m.engine.head_off[1], m.engine.head_off[2]
    assert torch.equal(m.theta[lo1:hi1], before[lo1:hi1])                         # the other heads receive no gradient and do not move (gpm.py:70)
    act = torch.zeros_like(before, dtype=torch.bool)
    for lo_, hi_ in m.active_ranges():
        act[lo_:hi_] = True
    assert torch.allclose(m.theta[act], (before - 0.01 * gsnap)[act], rtol=1e-6, atol=1e-8) and torch.equal(m.theta[~act], before[~act])

    def sync():
        for o in orcs.values():
            with torch.no_grad():
                for n in o.p:
                    o.p[n].copy_(m.engine.param_view(n).cpu())
                for t in range(3):
                    o.heads[t].copy_(m.engine.head_view(t).cpu())
    sync()
    # task boundary: bases from 125 samples
    x_all = pool[:150]
    loader = [{"image": x_all[i:i + 40]} for i in range(0, 150, 40)]
    torch.manual_seed(900)
    m.after_task(0, None, loader, None)
    torch.manual_seed(900)
    sel = torch.randperm(150)[:125]
    for o in orcs.values():
        o.after_task(x_all[sel])
    ranks = [f.shape[1] for f in m.feature_list]
    print("ranks", ranks, "fp32 oracle", [f.shape[1] for f in orcs["fp32"].feature_list], "bf16 oracle", [f.shape[1] for f in orcs["bf16"].feature_list],
          "reference", list(g["t0/rank"]))
    prng = np.random.default_rng(77)
    for i, (f, fo) in enumerate(zip(m.feature_list, orcs["fp32"].feature_list)):
        assert abs(f.shape[1] - int(g["t0/rank"][i])) <= 3, (i, f.shape[1], int(g["t0/rank"][i]))
        assert np.allclose(f.T @ f, np.eye(f.shape[1]), atol=1e-8)                     # orthonormal columns
        v = prng.standard_normal(f.shape[0])
        assert rel_l2(f @ (f.T @ v), g[f"t0/proj_probe/{i}"]) <= 5e-2, (i, rel_l2(f @ (f.T @ v), g[f"t0/proj_probe/{i}"]))
    # task 1 from the CUDA path's own bases on both sides (the comparison is the projected step, not the basis again)
    for o in orcs.values():
        o.feature_list = [f.copy() for f in m.feature_list]
        o.before_task(1)
    m.before_task(1, None, None, None)
    assert m._bn_frozen and not m.network.backbone.bn1.weight.requires_grad and m.network.backbone.conv1.weight.requires_grad
    for s in range(2):
        x, y = pool[20 + 16 * s:36 + 16 * s], torch.from_numpy(rng.integers(10, 20, (B,)).astype(np.int64))
        got = _step_check(m, orcs, x, y, g, f"t1s{s}")
        for i, name in enumerate(m.layers):                 # g' U = 0: the projected gradient has no component inside the stored subspace
            gw = got[name + ".weight"].double().reshape(got[name + ".weight"].shape[0], -1)
            U = torch.from_numpy(m.feature_list[i])
            assert float((gw @ U).norm() / gw.norm()) <= 1e-4, (name, float((gw @ U).norm() / gw.norm()))
        opt = FlatSGD(m.get_parameters(None), lr=0.01, model=m)
        before = m.engine.param_view("bn3.weight").clone()
        opt.step()
        assert torch.equal(before, m.engine.param_view("bn3.weight"))          # BN affine frozen after task 0 (gpm.py:126-129)
        sync()
    x_all = pool[10:160]
    torch.manual_seed(901)
    m.after_task(1, None, [{"image": x_all[i:i + 50]} for i in range(0, 150, 50)], None)
    r1 = [f.shape[1] for f in m.feature_list]
    print("ranks after task 1", r1, "reference", list(g["t1/rank"]))
    assert all(a >= b for a, b in zip(r1, ranks)) and all(abs(a - int(b)) <= 6 for a, b in zip(r1, g["t1/rank"]))
    # inference: task-aware and task-agnostic (gpm.py:85-111)
    xt, yt = pool[:32], torch.from_numpy(rng.integers(0, 20, (32,)).astype(np.int64))
    pa, _ = m.inference({"image": xt, "label": yt}, task_id=1)
    pg, _ = m.inference({"image": xt, "label": yt})
    feat = port.alexnet_forward({k: v.detach() for k, v in orcs["bf16"].p.items()}, xt, None, "bf16")
    lg = [feat @ h.detach().T for h in orcs["bf16"].heads]
    assert float((pa.cpu() == lg[1].argmax(1) + 10).float().mean()) >= 0.9 and float((pg.cpu() == torch.cat(lg, 1).argmax(1)).float().mean()) >= 0.9


def test_gpm_train_mode_dropout_step_vs_oracle_with_the_same_masks():
    """Train mode: the keep masks of the step are a pure function of (seed, step, layer, element) — read back through `lc_nn_dropout_mask` and fed to the
    oracle, so the dropout path (mask, 1/(1-p) scale, its backward) is compared exactly like everything else."""
    from libcontinual_b200._lib import check
    p, heads = synth_alexnet_state(5050)
    m = make_model(p, heads, max_batch=B)
    m.train()
    m.before_task(0, None, None, None)
    rng, pool = _pool()
    x, y = pool[40:40 + B], torch.from_numpy(rng.integers(0, 10, (B,)).astype(np.int64))
    pred, acc, loss = m.observe({"image": x, "label": y})
    torch.cuda.synchronize()
    eng = m.engine
    masks = []
    for i, L in enumerate(eng.layers):
        n = B * L.Ho * L.Ho * L.cout
        keep = torch.empty(n, dtype=torch.uint8, device="cuda")
        check(eng.lib.lc_nn_dropout_mask(eng.rng.data_ptr(), i, L.p, n, keep.data_ptr(), torch.cuda.current_stream().cuda_stream))
        k = keep.view(B, L.Ho, L.Ho, L.cout).permute(0, 3, 1, 2).contiguous().cpu().bool() if i < 3 else keep.view(B, L.cout).cpu().bool()
        assert abs(float(k.float().mean()) - (1 - L.p)) < 0.02
        masks.append(k)
    orc = port.GPMOracle(p, heads, 10, 10, gemm_mode="bf16")
    orc.before_task(0)
    po, ao, lo, go = orc.step(x, y, masks=masks, apply_update=False)
    got = grads_of(m)
    errs = {k: rel_l2(got[k], go[k]) for k in go}
    print(f"train-mode step: loss {float(loss):.5f} vs oracle {float(lo):.5f}; gradient rel-L2 max {max(errs.values()):.3e} ({max(errs, key=errs.get)})")
    assert abs(float(loss) - float(lo)) <= 5e-3 and max(errs.values()) <= 6e-2
    # a second step draws new masks
    step0 = int(eng.rng[1])
    m.observe({"image": x, "label": y})
    assert int(eng.rng[1]) == step0 + 1


def test_gpm_graphed_step_equals_eager():
    from libcontinual_b200.optim import FlatSGD
    from libcontinual_b200.trainer import GraphedFlatStep
    outs = []
    rng, pool = _pool()
    ys = [torch.from_numpy(rng.integers(10, 20, (64,)).astype(np.int64)) for _ in range(2)]
    for graphed in (False, True):
        p, heads = synth_alexnet_state(6060)
        m = make_model(p, heads, max_batch=64)
        m.train()
        m.before_task(0, None, None, None)
        prng = np.random.default_rng(1)
        m.feature_list = [np.linalg.qr(prng.standard_normal((d, r)))[0] for d, r in ((48, 20), (576, 100), (512, 90), (1024, 60), (2048, 80))]
        m.before_task(1, None, None, None)
        opt = FlatSGD(m.get_parameters(None), lr=0.01, model=m)
        step = GraphedFlatStep(m, opt, 64, img=32) if graphed else None
        m.engine.rng[1] = 100                               # both arms draw the same dropout masks
        for s in range(2):
            x, y = pool[64 * s:64 * s + 64], ys[s]
            if graphed:
                step.run(x.cuda(), y.cuda())
            else:
                opt.zero_grad()
                m.observe({"image": x, "label": y})
                opt.step()
        torch.cuda.synchronize()
        outs.append((m.theta.clone(), float(m.scal[0])))
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]
