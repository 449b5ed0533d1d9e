"""GPU: every CUDA kernel family, called through the C ABI (ctypes), against plain PyTorch fp32 CPU ops on the same
seeded inputs.  fp32 kernels with a different summation order: rtol/atol 1e-4 unless stated."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

RT, AT = 1e-4, 1e-4


@pytest.fixture(scope="module")
def lib():
    from libcontinual_b200 import _lib
    return _lib.load()


_KEEP = []


@pytest.fixture(autouse=True)
def _keepalive():
    """Raw pointers are handed to the C ABI: every device temporary must outlive the (asynchronous) call."""
    _KEEP.clear()
    yield
    torch.cuda.synchronize()
    _KEEP.clear()


def dev(t):
    d = t.cuda().contiguous()
    _KEEP.append(d)
    return d


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


def st():
    return torch.cuda.current_stream().cuda_stream


def P(t):
    return None if t is None else t.data_ptr()


def close(a, b, rt=RT, at=AT, what=""):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    err = (a - b).abs().max().item()
    assert torch.allclose(a, b, rtol=rt, atol=at), f"{what}: max|d|={err:.3e} ref max={b.abs().max().item():.3e}"


CONV_SHAPES = [  # cin, cout, width_out, stride
    (16, 16, 32, 1), (16, 32, 16, 2), (32, 32, 16, 1), (32, 64, 8, 2), (64, 64, 8, 1)]


@pytest.mark.parametrize("cin,cout,wo,stride", CONV_SHAPES)
@pytest.mark.parametrize("B", [3, 8])
def test_conv3x3_forward_stats_prologue(lib, cin, cout, wo, stride, B):
    g = torch.Generator().manual_seed(cin * 1000 + cout + B)
    wi = wo * stride
    x = torch.randn(B, cin, wi, wi, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    ps, psh = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    rm, rv = torch.randn(cout, generator=g), torch.rand(cout, generator=g) + 0.5
    for prologue in (False, True):
        xin = F.relu(x * ps.view(1, -1, 1, 1) + psh.view(1, -1, 1, 1)) if prologue else x
        ref = F.conv2d(xin, w, None, stride, 1)
        rm_ref, rv_ref = rm.clone(), rv.clone()
        bn_ref = F.batch_norm(ref, rm_ref, rv_ref, gamma, beta, True, 0.1, 1e-5)
        out = torch.empty(B, wo, wo, cout, device="cuda")
        scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, cin, cout, wo)), device="cuda")
        stat = torch.zeros(4 * cout, device="cuda")
        rstat = dev(torch.cat([rm, rv]))
        rc = lib.lc_conv3x3(P(dev(nhwc(x))), P(dev(w)), P(out), B, cin, cout, wo, stride, 0, 0, P(dev(ps)) if prologue else None,
                            P(dev(psh)) if prologue else None, None, P(dev(gamma)), P(dev(beta)), P(rstat), P(stat), P(scratch), st())
        assert rc == 0
        torch.cuda.synchronize()
        close(nchw(out), ref, what="conv out")
        mean, var = ref.mean((0, 2, 3)), ref.var((0, 2, 3), unbiased=False)
        close(stat[2 * cout:3 * cout], mean, 1e-4, 1e-5, "mean")
        close(stat[3 * cout:], 1 / torch.sqrt(var + 1e-5), 1e-4, 1e-5, "invstd")
        close(rstat[:cout], rm_ref, 1e-4, 1e-5, "running_mean"); close(rstat[cout:], rv_ref, 1e-4, 1e-5, "running_var")
        # affine form reproduces batch_norm
        y = nchw(out).cpu() * stat[:cout].cpu().view(1, -1, 1, 1) + stat[cout:2 * cout].cpu().view(1, -1, 1, 1)
        close(y, bn_ref, 1e-4, 2e-4, "bn affine")


def test_conv3x3_stem_nchw(lib):
    g = torch.Generator().manual_seed(5)
    B = 4
    x = torch.randn(B, 3, 32, 32, generator=g)
    w = torch.randn(16, 3, 3, 3, generator=g) * 0.2
    out = torch.empty(B, 32, 32, 16, device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, 3, 16, 32)), device="cuda")
    assert lib.lc_conv3x3(P(dev(x)), P(dev(w)), P(out), B, 3, 16, 32, 1, 0, 1, None, None, None, None, None, None, None, P(scratch), st()) == 0
    close(nchw(out), F.conv2d(x, w, None, 1, 1), what="stem")
    dy = torch.randn(B, 16, 32, 32, generator=g)
    dw = torch.empty(16, 3, 3, 3, device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, 3, 16, 32)), device="cuda")
    assert lib.lc_conv3x3_wgrad(P(dev(x)), P(dev(nhwc(dy))), P(dw), B, 3, 16, 32, 1, 1, None, None, P(scratch), st()) == 0
    ref = torch.nn.grad.conv2d_weight(x, w.shape, dy, stride=1, padding=1)
    close(dw, ref, 1e-4, 1e-3, "stem wgrad")


@pytest.mark.parametrize("cin,cout,wo,stride", CONV_SHAPES)
def test_conv3x3_dgrad_and_wgrad(lib, cin, cout, wo, stride):
    B = 5
    g = torch.Generator().manual_seed(cin + 7 * cout)
    wi = wo * stride
    x = torch.randn(B, cin, wi, wi, generator=g, requires_grad=True)
    w = (torch.randn(cout, cin, 3, 3, generator=g) * 0.1).requires_grad_(True)
    ps, psh = torch.rand(cin, generator=g) + 0.5, torch.randn(cin, generator=g) * 0.3
    dy = torch.randn(B, cout, wo, wo, generator=g)
    addend = torch.randn(B, cin, wi, wi, generator=g)
    y = F.conv2d(x, w, None, stride, 1)
    dx_ref, dw_ref = torch.autograd.grad(y, [x, w], dy)
    dx = torch.empty(B, wi, wi, cin, device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, cin, cout, wi)), device="cuda")
    add_d = dev(nhwc(addend)) if stride == 1 else None
    assert lib.lc_conv3x3(P(dev(nhwc(dy))), P(dev(w.detach())), P(dx), B, cin, cout, wo, stride, 1, 0, None, None, P(add_d), None, None, None, None,
                          P(scratch), st()) == 0
    close(nchw(dx), dx_ref + (addend if stride == 1 else 0), what="dgrad")
    # wgrad without and with the BN+ReLU prologue on the input
    for prologue in (False, True):
        xin = x.detach()
        if prologue:
            xr = F.relu(xin * ps.view(1, -1, 1, 1) + psh.view(1, -1, 1, 1))
            dw_r = torch.nn.grad.conv2d_weight(xr, w.shape, dy, stride=stride, padding=1)
        else:
            dw_r = dw_ref
        dw = torch.empty(cout, cin, 3, 3, device="cuda")
        scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, cin, cout, wo)), device="cuda")
        assert lib.lc_conv3x3_wgrad(P(dev(nhwc(xin))), P(dev(nhwc(dy))), P(dw), B, cin, cout, wo, stride, 0, P(dev(ps)) if prologue else None,
                                    P(dev(psh)) if prologue else None, P(scratch), st()) == 0
        close(dw, dw_r, 1e-4, 2e-3, f"wgrad prologue={prologue}")


@pytest.mark.parametrize("cin,cout,wo", [(16, 32, 16), (32, 64, 8)])
@pytest.mark.parametrize("B", [5, 16])
def test_conv1x1_stride2(lib, cin, cout, wo, B):
    g = torch.Generator().manual_seed(cin + B)
    x = torch.randn(B, cin, 2 * wo, 2 * wo, generator=g, requires_grad=True)
    w = (torch.randn(cout, cin, 1, 1, generator=g) * 0.2).requires_grad_(True)
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    y = F.conv2d(x, w, None, 2, 0)
    out = torch.empty(B, wo, wo, cout, device="cuda")
    scratch = torch.zeros(200000, device="cuda")
    stat = torch.zeros(4 * cout, device="cuda")
    assert lib.lc_conv1x1s2(P(dev(nhwc(x.detach()))), P(dev(w.detach())), P(out), B, cin, cout, wo, 0, P(dev(gamma)), P(dev(beta)), None, P(stat), P(scratch), st()) == 0
    close(nchw(out), y, what="1x1 fwd")
    close(stat[2 * cout:3 * cout], y.mean((0, 2, 3)), 1e-4, 1e-5, "1x1 mean")
    close(stat[3 * cout:], 1 / torch.sqrt(y.var((0, 2, 3), unbiased=False) + 1e-5), 1e-4, 1e-5, "1x1 invstd")
    dy = torch.randn(B, cout, wo, wo, generator=g)
    dx_ref, dw_ref = torch.autograd.grad(y, [x, w], dy)
    base = torch.randn(B, cin, 2 * wo, 2 * wo, generator=g)
    acc = dev(nhwc(base))
    assert lib.lc_conv1x1s2(P(dev(nhwc(dy))), P(dev(w.detach())), P(acc), B, cin, cout, wo, 1, None, None, None, None, P(scratch), st()) == 0
    close(nchw(acc), base + dx_ref, what="1x1 dgrad accumulate")
    dw = torch.empty(cout, cin, device="cuda")
    scratch = torch.zeros(200000, device="cuda")
    assert lib.lc_conv1x1s2(P(dev(nhwc(x.detach()))), P(dev(nhwc(dy))), P(dw), B, cin, cout, wo, 2, None, None, None, None, P(scratch), st()) == 0
    close(dw, dw_ref.view(cout, cin), 1e-4, 1e-3, "1x1 wgrad")


@pytest.mark.parametrize("C,wo", [(16, 32), (32, 16), (64, 8)])
@pytest.mark.parametrize("mask_mode", [0, 1, 2])
def test_bn_act_forward_backward(lib, C, wo, mask_mode):
    B = 6
    g = torch.Generator().manual_seed(C + mask_mode)
    y = torch.randn(B, C, wo, wo, generator=g, requires_grad=True)
    gamma = (torch.rand(C, generator=g) + 0.5).requires_grad_(True)
    beta = (torch.randn(C, generator=g) * 0.5).requires_grad_(True)
    res = torch.randn(B, C, wo, wo, generator=g)
    gin = torch.randn(B, C, wo, wo, generator=g)
    bn = F.batch_norm(y, None, None, gamma, beta, True, 0.1, 1e-5)
    if mask_mode == 0:
        out = bn
    elif mask_mode == 1:
        out = F.relu(bn + res)
    else:
        out = F.relu(bn)
    dy_ref, dg_ref, db_ref = torch.autograd.grad(out, [y, gamma, beta], gin)
    mean, var = y.detach().mean((0, 2, 3)), y.detach().var((0, 2, 3), unbiased=False)
    invstd = 1 / torch.sqrt(var + 1e-5)
    scale = gamma.detach() * invstd
    shift = beta.detach() - mean * scale
    stat = dev(torch.cat([scale, shift, mean, invstd]))
    npix = B * wo * wo
    yd = dev(nhwc(y.detach()))
    # forward elementwise kernel
    o = torch.empty(B, wo, wo, C, device="cuda")
    assert lib.lc_bn_act_forward(P(yd), P(stat), stat.data_ptr() + 4 * C, P(dev(nhwc(res))) if mask_mode == 1 else None, None, None, P(o), npix, C, st()) == 0
    fwd_ref = F.relu(bn + res) if mask_mode == 1 else F.relu(bn)
    close(nchw(o), fwd_ref, 1e-4, 1e-4, "bn_act fwd")
    # backward
    dy = torch.empty(B, wo, wo, C, device="cuda")
    gout = torch.empty(B, wo, wo, C, device="cuda")
    dgam, dbet = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    scratch = torch.zeros(592 * 2 * C + 3 * C + 128, device="cuda")
    mask_src = o if mask_mode == 1 else None
    assert lib.lc_bn_backward(P(dev(nhwc(gin))), P(mask_src), mask_mode, P(yd), P(stat), P(dy), P(gout), P(dgam), P(dbet), npix, C, P(scratch), st()) == 0
    close(nchw(dy), dy_ref, 1e-4, 1e-4, "bn bwd dy")
    close(dgam, dg_ref, 1e-4, 1e-3, "dgamma"); close(dbet, db_ref, 1e-4, 1e-3, "dbeta")
    if mask_mode == 1:
        close(nchw(gout), gin * (fwd_ref > 0), 0, 0, "masked g")


def test_bn_act_forward_downsample_residual(lib):
    B, C, wo = 4, 32, 16
    g = torch.Generator().manual_seed(11)
    y, r = torch.randn(B, wo, wo, C, generator=g), torch.randn(B, wo, wo, C, generator=g)
    a = torch.randn(4, C, generator=g)
    o = torch.empty(B, wo, wo, C, device="cuda")
    assert lib.lc_bn_act_forward(P(dev(y)), P(dev(a[0])), P(dev(a[1])), P(dev(r)), P(dev(a[2])), P(dev(a[3])), P(o), B * wo * wo, C, st()) == 0
    close(o, F.relu(y * a[0] + a[1] + r * a[2] + a[3]), 1e-5, 1e-5)


@pytest.mark.parametrize("B,ncls,ce_lo,kd_n", [(8, 10, 0, 0), (32, 20, 10, 0), (128, 15, 0, 10), (33, 100, 80, 80), (5, 100, 0, 50)])
def test_head_and_loss(lib, B, ncls, ce_lo, kd_n):
    g = torch.Generator().manual_seed(B + ncls)
    cap, C, HW = 100, 64, 64
    act = torch.rand(B, HW, C, generator=g)
    W = (torch.randn(ncls, C, generator=g) * 0.3).requires_grad_(True)
    bias = (torch.randn(ncls, generator=g) * 0.1).requires_grad_(True)
    act_r = act.clone().requires_grad_(True)
    y = torch.randint(ce_lo, ncls, (B,), generator=g)
    teacher = torch.randn(B, cap, generator=g)
    kd_w, T = 3.0, 2.0
    feat_ref = act_r.mean(1)
    logits_ref = F.linear(feat_ref, W, bias)
    loss_ref = F.cross_entropy(logits_ref[:, ce_lo:], y - ce_lo)
    if kd_n:
        kd = -(torch.softmax(teacher[:, :kd_n] / T, 1) * torch.log_softmax(logits_ref[:, :kd_n] / T, 1)).sum() / B
        loss_ref = loss_ref + kd_w * kd
    dW_ref, db_ref, dact_ref = torch.autograd.grad(loss_ref, [W, bias, act_r])
    feat = torch.empty(B, C, device="cuda"); logits = torch.zeros(B, cap, device="cuda")
    Wd, bd, actd = dev(W.detach()), dev(bias.detach()), dev(act)
    assert lib.lc_head_forward(P(actd), B, HW, C, P(Wd), P(bd), ncls, P(feat), P(logits), cap, st()) == 0
    close(feat, feat_ref, 1e-5, 1e-6, "feat"); close(logits[:, :ncls], logits_ref, 1e-5, 1e-5, "logits")
    dl = torch.full((B, cap), 7.0, device="cuda"); pred = torch.zeros(B, dtype=torch.int64, device="cuda"); scal = torch.zeros(8, device="cuda")
    assert lib.lc_loss_ce_kd(P(logits), cap, P(dev(teacher)) if kd_n else None, cap, P(dev(y)), B, ce_lo, ncls, kd_n, kd_w, T, ncls, P(dl), P(pred), P(scal), st()) == 0
    close(scal[0], loss_ref, 1e-5, 1e-6, "loss")
    assert torch.equal(pred.cpu(), logits_ref.argmax(1))
    assert int(scal[1].item()) == int((logits_ref.argmax(1) == y).sum())
    assert float(dl[:, ncls:].abs().max()) == 0.0 if ncls < cap else True
    dWd, dbd = torch.zeros(ncls, C, device="cuda"), torch.zeros(ncls, device="cuda")
    dfeat, gact = torch.empty(B, C, device="cuda"), torch.empty(B, HW, C, device="cuda")
    assert lib.lc_head_backward(P(dl), cap, P(feat), P(Wd), ncls, B, C, P(dWd), P(dbd), P(dfeat), P(gact), HW, st()) == 0
    close(dWd, dW_ref, 1e-4, 1e-6, "dW"); close(dbd, db_ref, 1e-4, 1e-6, "db"); close(gact, dact_ref, 1e-4, 1e-7, "gact")


@pytest.mark.parametrize("n", [466256 + 6500, 1001, 4])
def test_flat_ops(lib, n):
    g = torch.Generator().manual_seed(n)
    n4 = (n + 3) // 4 * 4
    theta, ref, fisher, grad = (torch.randn(n4, generator=g) for _ in range(4))
    fisher = fisher.abs()
    lam = 1000.0
    hp = dev(torch.tensor([lam]))
    scal = torch.zeros(8, device="cuda"); scal[0] = 0.25
    scratch = torch.zeros(2 * 296 + 8, device="cuda"); counter = torch.zeros(4, dtype=torch.int32, device="cuda")
    gd = dev(grad)
    for rep in range(2):      # second call checks that the election counter resets itself
        gd.copy_(grad); scal[0] = 0.25
        assert lib.lc_ewc_penalty_grad(P(dev(theta)), P(dev(ref)), P(dev(fisher)), P(gd), n4, P(hp), P(scratch), P(counter), P(scal), st()) == 0
        pen = (fisher.double() * (theta.double() - ref.double()) ** 2).sum() / 2
        close(gd, grad + lam * fisher * (theta - ref), 1e-5, 1e-4, "ewc grad")
        assert abs(scal[4].item() - pen.item()) <= 1e-5 * pen.item()
        assert abs(scal[0].item() - (0.25 + lam * pen.item())) <= 1e-5 * lam * pen.item()
    # fisher accumulate / merge
    fd = dev(fisher)
    assert lib.lc_fisher_accumulate(P(fd), P(dev(grad)), n4, 32.0, st()) == 0
    close(fd, fisher + grad * grad * 32.0, 1e-6, 1e-6, "fisher acc")
    old = torch.rand(n4, generator=g)
    assert lib.lc_fisher_merge(P(fd), P(dev(old)), n4, 5024.0, 0.75, st()) == 0
    close(fd, 0.75 * old + 0.25 * (fisher + grad * grad * 32.0) / 5024.0, 1e-5, 1e-7, "fisher merge")
    # SGD momentum: two steps against torch.optim.SGD
    p = torch.randn(n4, generator=g); gr = torch.randn(n4, generator=g)
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.SGD([pt], lr=0.1, momentum=0.9, weight_decay=5e-4)
    pd, md = dev(p), torch.zeros(n4, device="cuda")
    hp3 = dev(torch.tensor([0.1, 0.9, 5e-4, 0.0]))
    for k in range(2):
        pt.grad = gr.clone() * (k + 1)
        opt.step()
        assert lib.lc_sgd_momentum(P(pd), P(dev(gr * (k + 1))), P(md), n4, P(hp3), st()) == 0
    close(pd, pt, 1e-6, 1e-6, "sgd")
    # Adam
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=1.875e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0)
    pd, md, vd = dev(p), torch.zeros(n4, device="cuda"), torch.zeros(n4, device="cuda")
    for k in range(1, 3):
        pt.grad = gr.clone() * k
        opt.step()
        hpa = dev(torch.tensor([1.875e-3, 0.9, 0.999, 1e-8, 0.0, 1 - 0.9 ** k, 1 - 0.999 ** k]))
        assert lib.lc_adam(P(pd), P(dev(gr * k)), P(md), P(vd), n4, P(hpa), st()) == 0
    close(pd, pt, 1e-5, 1e-6, "adam")
    # clip_grad_norm_
    gt = (gr * 3).clone()
    gcl = dev(gt)
    nrm = torch.zeros(1, device="cuda")
    assert lib.lc_clip_grad_norm(P(gcl), n4, 1.0, P(scratch), P(nrm), st()) == 0
    pr = gt.clone().requires_grad_(True); pr.grad = gt.clone()
    total = torch.nn.utils.clip_grad_norm_([pr], 1.0)
    close(nrm, total, 1e-5, 1e-6, "grad norm"); close(gcl, pr.grad, 1e-5, 1e-7, "clipped grad")
