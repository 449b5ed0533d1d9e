"""GPU: parity at the BENCHMARKED configuration — batch 128, `precision="tc"` (and the exact fp32 mode beside it) — for the two ResNet32 workloads
`bench.py` times (iCaRL task 1 with the KD teacher live, icarl.py:197-221; EWC task 1 with the penalty live, ewc.py:82-108), against
`oracle.port.ResNetMethodOracle` on the same seeded inputs, plus the tensor-core conv / weight-gradient kernels alone at B = 128 and the ViT methods
at B = 16.

Tolerances are FIXED numbers; the measured values of every quantity are printed and written to `gpurun_out/parity_b128.json` (committed copy:
`profiles/r2_parity_b128.json`).  Measured on B200 (round 2) in brackets:

  quantity                       fp32 mode vs fp32 oracle            tc mode vs fp32 oracle
  loss |d|                       <= 1e-5 + 1e-5 rel   [0]            <= 1e-3                [6e-5]
  logits, features (rel-L2)      <= 1e-4              [< 1e-6]       <= 5e-3                [9e-4 .. 1e-3]
  gradient, whole arena (rel-L2) <= 2e-2              [2e-3 .. 5e-3] <= 2.5e-1              [0.089 (EWC) .. 0.141 (iCaRL)]
  gradient, per tensor (rel-L2)  median <= 2e-2, max <= 5e-2         median <= 2.5e-1, max <= 6e-1
                                 [median 1e-3..5e-3, max 3e-3..1.2e-2]   [median 0.127..0.141, max 0.29..0.37]
  tc mode only: every gradient figure <= 1.25 x (+1e-2) the figure the REFERENCE'S OWN GPU ARITHMETIC shows on the same step — the oracle's op sequence
  as eager PyTorch on cuda:0 with cuDNN TF32 convolutions (PyTorch's default, never changed by the reference) against the same ops in strict fp32
  [whole arena 0.0891 / 0.1390, median 0.129 / 0.141, max 0.276 / 0.359: the same numbers as ours to two digits]
  pred / #correct                exact                               >= 97 % of the batch identical [127 / 128 .. 128 / 128: near-tied
                                                                     logits of a randomly initialised head under operand rounding]

Why gradients move by 10 % when logits move by 1e-3: the gradient of a randomly initialised BN + ReLU ResNet32 is discontinuous in its activations.  A
relative perturbation d of the activations flips the ReLU mask of a fraction ~d of the units, and every flip changes the gradient of all upstream
parameters; over 31 layers the gradient error grows like ~10 sqrt(d): fp32 round-off (6e-8) -> 2e-3 .. 5e-3 (the fp32 row above, and the fp32 CPU oracle
against an fp64 run of itself), TF32 operand rounding (2^-11) -> 0.1 .. 0.2.  That is a property of the function being differentiated, not of this
implementation, which is what the reference-GPU-arithmetic row demonstrates: cuDNN's TF32 path — what the reference itself runs on a GPU — sits exactly as
far from the reference's CPU path as this library's tensor-core mode does.
"""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import port
from tests.golden_util import synth_batch, synth_resnet_state
from tests.test_gpu_resnet import grads_of, load_head, rel_l2, sync_oracle_from

pytestmark = pytest.mark.gpu
B = 128
_REPORT = {}


def _record(key, val):
    _REPORT[key] = val
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_b128.json"), "w") as f:
            json.dump(_REPORT, f, indent=1, sort_keys=True)
    except OSError:
        pass


TOL = {"fp32": dict(loss_abs=1e-5, loss_rel=1e-5, logits=1e-4, feat=1e-4, g_med=2e-2, g_max=5e-2, g_all=2e-2, pred_frac=1.0),
       "tc": dict(loss_abs=1e-3, loss_rel=0.0, logits=5e-3, feat=5e-3, g_med=2.5e-1, g_max=6e-1, g_all=2.5e-1, pred_frac=0.97)}


def _reference_gpu_arithmetic(orc, x, y):
    """The reference's OWN GPU arithmetic on the same state and batch: the oracle's op sequence as eager PyTorch on cuda:0 with cuDNN TF32 convolutions
    allowed (PyTorch's default, which the reference never changes: SURVEY 2.3) against the same ops in strict fp32 — how far the reference's GPU path
    sits from its CPU path on this very step.  Returns whole-arena / median / max per-tensor gradient rel-L2 and |d loss|."""
    import copy
    dev = torch.device("cuda")
    out = {}
    for tf32 in (True, False):
        o = copy.copy(orc)
        o.p = {k: v.detach().to(dev).requires_grad_(True) for k, v in orc.p.items()}
        o.b = {k: v.to(dev).clone() for k, v in orc.b.items()}
        o.fc_w = orc.fc_w.detach().to(dev).requires_grad_(True); o.fc_b = orc.fc_b.detach().to(dev).requires_grad_(True)
        if orc.teacher is not None:
            tp, tb, tw, tbias = orc.teacher
            o.teacher = ({k: v.to(dev) for k, v in tp.items()}, {k: v.to(dev) for k, v in tb.items()}, tw.to(dev), tbias.to(dev))
        o.ref = {k: v.to(dev) for k, v in orc.ref.items()}; o.fisher = {k: v.to(dev) for k, v in orc.fisher.items()}
        saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf32, False
        try:
            _, _, l, g = o.step(x.to(dev), y.to(dev), apply_update=False)
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
        out[tf32] = (float(l), {k: v.detach().cpu() for k, v in g.items()})
    (l_tf, g_tf), (l_32, g_32) = out[True], out[False]
    errs = [rel_l2(g_tf[k], g_32[k]) for k in g_32]
    a = torch.cat([g_tf[k].reshape(-1).double() for k in g_32]); b = torch.cat([g_32[k].reshape(-1).double() for k in g_32])
    return {"grad_rel_l2_whole_arena": float((a - b).norm() / b.norm()), "grad_rel_l2_median": float(np.median(errs)), "grad_rel_l2_max": max(errs),
            "loss_abs_err": abs(l_tf - l_32)}


def _compare(tag, m, orc, x, y, precision, p_for_fwd, opt):
    """observe -> zero_grad -> backward (trainer.py:601-604) of the CUDA path vs one oracle step from identical state; returns the measured numbers.
    The caller finishes the step with `opt.step()`."""
    eng = m.engine
    pred, acc, loss = m.observe({"image": x, "label": y})
    opt.zero_grad()
    loss.backward()
    torch.cuda.synchronize()
    assert not eng.tensor_core_error()
    got = grads_of(m)
    n = eng.ncls
    logits = eng.logits[:B, :n].clone().cpu()
    feat = eng.features(B).clone().cpu()
    ob = {k: v.clone() for k, v in orc.b.items()}
    ref_gpu = _reference_gpu_arithmetic(orc, x, y) if precision == "tc" else None
    po, ao, lo, go = orc.step(x, y, apply_update=False)
    with torch.no_grad():
        ofeat = port.cifar_resnet_forward({k: v.detach() for k, v in p_for_fwd.items()}, ob, x, True)["features"]
        ologits = F.linear(ofeat, orc.fc_w.detach(), orc.fc_b.detach())
    errs = {k: rel_l2(got[k], go[k]) for k in go}
    flat_g = torch.cat([got[k].reshape(-1).double() for k in go]); flat_o = torch.cat([go[k].reshape(-1).double() for k in go])
    res = {"loss": float(loss), "oracle_loss": float(lo), "loss_abs_err": abs(float(loss) - float(lo)), "logits_rel_l2": rel_l2(logits, ologits),
           "features_rel_l2": rel_l2(feat, ofeat), "grad_rel_l2_median": float(np.median(list(errs.values()))), "grad_rel_l2_max": max(errs.values()),
           "grad_rel_l2_worst_tensor": max(errs, key=errs.get), "grad_rel_l2_whole_arena": float((flat_g - flat_o).norm() / flat_o.norm()),
           "pred_equal_frac": float((pred.cpu() == po).float().mean()), "correct": acc * B, "oracle_correct": ao * B}
    if ref_gpu is not None:
        res["reference_cudnn_tf32_vs_fp32_same_step"] = ref_gpu
    print(f"[{tag} / {precision}] " + ", ".join(f"{k}={v:.3e}" if isinstance(v, float) else f"{k}={v}" for k, v in res.items()))
    _record(f"{tag}/{precision}", res)
    t = TOL[precision]
    assert res["loss_abs_err"] <= t["loss_abs"] + t["loss_rel"] * abs(float(lo)), res
    assert res["logits_rel_l2"] <= t["logits"] and res["features_rel_l2"] <= t["feat"], res
    assert res["grad_rel_l2_median"] <= t["g_med"] and res["grad_rel_l2_max"] <= t["g_max"] and res["grad_rel_l2_whole_arena"] <= t["g_all"], res
    assert res["pred_equal_frac"] >= t["pred_frac"], res
    if ref_gpu is not None:                       # no further from the reference's CPU path than the reference's own GPU path is
        for k in ("grad_rel_l2_whole_arena", "grad_rel_l2_median", "grad_rel_l2_max"):
            assert res[k] <= 1.25 * ref_gpu[k] + 1e-2, (k, res[k], ref_gpu[k])
    return res


def _backbone(p, b, precision):
    import libcontinual_b200.model as M
    bb = M.cifar_resnet32(max_batch=B, precision=precision)
    bb.load_state_dict({**p, **b}, strict=True)
    return bb


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_icarl_task1_kd_b128_vs_oracle(precision):
    """BASELINE configs[1] as benched: 50 base classes, 5 new, batch 128, KD against the frozen teacher; two consecutive steps with the fused SGD in
    between (the second starts from a state that differs from the teacher's)."""
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    p, b, fc_w, fc_b = synth_resnet_state(4242, 100)
    bb = _backbone(p, b, precision)
    m = M.ICarl(bb, 64, 100, device=torch.device("cuda"), init_cls_num=50, inc_cls_num=5, task_num=11)
    load_head(m, fc_w, fc_b)
    m.before_task(0, None, None, None)
    m.train()
    opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=m.engine)
    orc = port.ResNetMethodOracle("icarl", p, b, fc_w, fc_b, init_cls=50, inc_cls=5)
    x, y = synth_batch(5000, B, 0, 50)
    _compare("icarl_t0s0", m, orc, x, y, precision, orc.p, opt)     # task 0: plain CE over 50 classes
    opt.step()
    sync_oracle_from(m, orc)
    m.snapshot_teacher(); m.cur_task_id += 1
    m.before_task(1, None, None, None)
    orc.snapshot_teacher(); orc.prev_cls = 50; orc.accu_cls = 55; orc.task_idx = 1; orc.reset_optimizer()
    opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=m.engine)
    for s in range(2):
        x, y = synth_batch(5100 + s, B, 0, 55)
        _compare(f"icarl_t1s{s}", m, orc, x, y, precision, orc.p, opt)
        assert float(m.engine.scal[3]) > 0.0                        # the KD term is live
        opt.step()
        sync_oracle_from(m, orc)


@pytest.mark.parametrize("precision", ["fp32", "tc"])
def test_ewc_task1_penalty_b128_vs_oracle(precision):
    """BASELINE configs[0] at the benched batch size: task 1, CE on the new slice + lamda/2 sum F (theta - theta*)^2 with lamda = 1000."""
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    p, b, fc_w, fc_b = synth_resnet_state(4343, 20)
    bb = _backbone(p, b, precision)
    m = M.EWC(bb, 64, 100, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10, lamda=1000.0)
    m.before_task(0, None, None, None)
    m.before_task(1, None, None, None)
    load_head(m, fc_w, fc_b)
    eng = m.engine
    rng = np.random.default_rng(9)
    m.ref_param = eng.params * torch.from_numpy(1 + 1e-2 * rng.standard_normal(eng.n_total)).float().cuda()
    m.fisher = torch.from_numpy(rng.uniform(0, 1e-3, eng.n_total).astype(np.float32)).cuda()
    m.fisher[eng.off_fc_w + 10 * 64:eng.off_fc_b] = 0
    m.fisher[eng.off_fc_b + 10:] = 0
    m.train()
    opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=eng)
    orc = port.ResNetMethodOracle("ewc", p, b, fc_w, fc_b, init_cls=10, inc_cls=10, lamda=1000.0)
    orc.task_idx = 1
    orc.ref = {"backbone." + n: eng.param_view(n, m.ref_param).cpu().clone() for n, _ in eng.layout}
    rw, rb = eng.fc_views(10, m.ref_param)
    orc.ref["classifier.weight"], orc.ref["classifier.bias"] = rw.cpu().clone(), rb.cpu().clone()
    orc.fisher = {"backbone." + n: eng.param_view(n, m.fisher).cpu().clone() for n, _ in eng.layout}
    fw, fb = eng.fc_views(10, m.fisher)
    orc.fisher["classifier.weight"], orc.fisher["classifier.bias"] = fw.cpu().clone(), fb.cpu().clone()
    for s in range(2):
        x, y = synth_batch(5200 + s, B, 10, 20)
        _compare(f"ewc_t1s{s}", m, orc, x, y, precision, orc.p, opt)
        assert float(eng.scal[4]) > 0.0                             # the penalty is live
        opt.step()
        sync_oracle_from(m, orc)


def test_graphed_step_matches_eager_b128_tc():
    """The path `bench.py` times (CUDA-graph replay, fused SGD inside) gives the same parameters as the eager plugin order, bit for bit, at B = 128 / tc."""
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    from libcontinual_b200.trainer import GraphedStep, train_step_eager
    outs = []
    for graphed in (False, True):
        p, b, fc_w, fc_b = synth_resnet_state(77, 100)
        bb = _backbone(p, b, "tc")
        m = M.ICarl(bb, 64, 100, device=torch.device("cuda"), init_cls_num=50, inc_cls_num=5, task_num=11)
        load_head(m, fc_w, fc_b)
        m.before_task(0, None, None, None)
        m.snapshot_teacher(); m.cur_task_id += 1
        m.before_task(1, None, None, None)
        m.train()
        opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=m.engine)
        step = GraphedStep(m, opt, B) if graphed else None
        for s in range(3):
            x, y = synth_batch(600 + s, B, 0, 55)
            if graphed:
                step.run(x.cuda(), y.cuda())
            else:
                train_step_eager(m, opt, {"image": x, "label": y})
        torch.cuda.synchronize()
        outs.append((m.engine.params.clone(), m.engine.rstat.clone(), float(m.engine.scal[0])))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]) and outs[0][2] == outs[1][2]


def test_graphed_step_prefetch_from_pinned_host_equals_device_batches():
    """`GraphedStep.run(pinned host batch)` with `.prefetch(next pinned host batch)` (the bench's e2e loop: double-buffered H2D on a copy stream) leaves the
    same parameters, bit for bit, as `run` on device-resident batches."""
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    from libcontinual_b200.trainer import GraphedStep
    outs = []
    batches = [synth_batch(700 + s, B, 0, 10) for s in range(4)]
    for mode in ("device", "prefetch"):
        p, b, fc_w, fc_b = synth_resnet_state(78, 100)
        bb = _backbone(p, b, "tc")
        m = M.EWC(bb, 64, 100, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10, lamda=1000.0)
        m.before_task(0, None, None, None)
        load_head(m, fc_w[:10], fc_b[:10])
        m.train()
        opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=m.engine)
        step = GraphedStep(m, opt, B)
        if mode == "device":
            for x, y in batches:
                step.run(x.cuda(), y.cuda())
        else:
            host = [(x.contiguous().pin_memory(), y.contiguous().pin_memory()) for x, y in batches]
            step.prefetch(*host[0])
            for j, (x, y) in enumerate(host):
                step.run(x, y)
                if j + 1 < len(host):
                    step.prefetch(*host[j + 1])
                step.loss().item()
        torch.cuda.synchronize()
        outs.append((m.engine.params.clone(), float(m.engine.scal[0])))
    assert torch.equal(outs[0][0], outs[1][0]) and outs[0][1] == outs[1][1]


# ---- the tensor-core kernels alone at the benched batch size ---------------------------------------------------------------------------
@pytest.mark.parametrize("c,w", [(16, 32), (32, 16), (64, 8)])
def test_conv3x3_tc_forward_b128(c, w):
    """The multi-tile grid of the bench (289 / 81 / 25 CTAs): max|err| <= 1e-3 * max|ref| against fp32 (TF32 operands; measured 3e-4), <= 5e-4 * max|ref|
    against an fp32 conv of TF32-rounded operands (measured 3e-5 .. 1.1e-4: summation order, plus the fused BN+ReLU prologue's FMA landing on the other
    side of a TF32 rounding point for a few operands), BN statistics from the epilogue within 1e-2 relative."""
    from tests.test_gpu_kernels import P, dev, nchw, nhwc, st
    from tests.test_gpu_tensorcore import tf32_round
    from libcontinual_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(1000 + c)
    x = torch.randn(B, c, w, w, generator=g)
    wt = torch.randn(c, c, 3, 3, generator=g) * 0.1
    ps, psh = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    xin = F.relu(x * ps.view(1, -1, 1, 1) + psh.view(1, -1, 1, 1))
    ref = F.conv2d(xin, wt, None, 1, 1)
    out = torch.full((B, w, w, c), float("nan"), device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_tc_scratch_floats(B, c, w)), device="cuda")
    stat = torch.zeros(4 * c, device="cuda")
    assert lib.lc_conv3x3_tc(P(dev(nhwc(x))), P(dev(wt)), P(out), B, c, w, 0, P(dev(ps)), P(dev(psh)), None, P(dev(gamma)), P(dev(beta)), None, P(stat),
                             P(scratch), st()) == 0
    torch.cuda.synchronize()
    assert int(scratch.view(torch.int32)[8]) == 0, "tensor-core barrier timed out"
    got = nchw(out).cpu()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    e_rn = (got - F.conv2d(tf32_round(xin), tf32_round(wt), None, 1, 1)).abs().max().item()
    _record(f"conv3x3_tc_fwd/c{c}w{w}", {"max_abs_err_over_max_ref": err / scale, "vs_tf32_rounded_operands": e_rn / scale})
    assert err <= 1e-3 * scale and e_rn <= 5e-4 * scale, (err, e_rn, scale)
    mean, var = ref.mean((0, 2, 3)), ref.var((0, 2, 3), unbiased=False)
    assert torch.allclose(stat[2 * c:3 * c].cpu(), mean, rtol=1e-2, atol=2e-3)
    assert torch.allclose(stat[3 * c:].cpu(), 1 / torch.sqrt(var + 1e-5), rtol=1e-2, atol=1e-3)


@pytest.mark.parametrize("c,w", [(16, 32), (32, 16), (64, 8)])
def test_conv3x3_tc_wgrad_b128(c, w):
    """Split-K over 296 / 148 / 50 CTAs as in the bench.  BF16 operands: |err| <= 5e-3 * max|ref| against fp32 (measured 2.6e-3 .. 2.8e-3; whole-tensor
    rel-L2 2.3e-3), <= 2e-4 against an fp32 contraction of the same BF16-rounded operands (measured <= 4e-5)."""
    from tests.test_gpu_kernels import P, dev, nhwc, st
    from libcontinual_b200 import _lib
    lib = _lib.load()
    g = torch.Generator().manual_seed(2000 + c)
    x = torch.randn(B, c, w, w, generator=g)
    dy = torch.randn(B, c, w, w, generator=g)
    ps, psh = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    xin = F.relu(x * ps.view(1, -1, 1, 1) + psh.view(1, -1, 1, 1))
    ref = torch.nn.grad.conv2d_weight(xin, (c, c, 3, 3), dy, stride=1, padding=1)
    ref_bf = torch.nn.grad.conv2d_weight(xin.bfloat16().float(), (c, c, 3, 3), dy.bfloat16().float(), stride=1, padding=1)
    dw = torch.full((c, c, 3, 3), float("nan"), device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, c, c, w)), device="cuda")
    assert lib.lc_conv3x3_wgrad_tc(P(dev(nhwc(x))), P(dev(nhwc(dy))), P(dw), B, c, w, P(dev(ps)), P(dev(psh)), P(scratch), st()) == 0
    torch.cuda.synchronize()
    assert int(scratch.view(torch.int32)[8]) == 0, "tensor-core barrier timed out"
    got = dw.cpu()
    scale = ref.abs().max().item()
    err, err_bf = (got - ref).abs().max().item(), (got - ref_bf).abs().max().item()
    _record(f"wgrad3x3_tc/c{c}w{w}", {"max_abs_err_over_max_ref": err / scale, "vs_bf16_rounded_operands": err_bf / scale,
                                      "rel_l2": rel_l2(got, ref)})
    assert err <= 5e-3 * scale and err_bf <= 2e-4 * scale, (err, err_bf, scale)
