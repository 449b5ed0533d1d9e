"""GPU: tcgen05 (TF32 operands, fp32 accumulation in TMEM) 3x3 convolution through the C ABI against PyTorch fp32 CPU conv.

Tolerance: TF32 keeps 10 mantissa bits of each operand (relative 2^-11 .. 2^-10 per product); outputs are sums of K = 9*C products
-> |err| <= 2e-3 * (|x| (*) |w|) elementwise is a hard bound; we assert max|err| <= 4e-3 * max|ref| and report the error against an
fp32 conv of TF32-truncated operands, which isolates the summation-order part (expected ~1e-6)."""
import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_kernels import P, close, dev, lib, nchw, nhwc, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def tf32_trunc(t):
    return (t.view(torch.int32) & ~0x1FFF).view(torch.float32)


def tf32_round(t):
    i = t.view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


@pytest.mark.parametrize("c,w", [(16, 32), (32, 16), (64, 8)])
@pytest.mark.parametrize("B", [1, 3, 8, 37, 200])      # 37: ragged tile ranges; 200: more CTAs than SMs for the persistent stage-1 kernel (conv_tcp.cuh)
def test_conv3x3_tc_forward(lib, c, w, B):
    g = torch.Generator().manual_seed(100 * c + B)
    x = torch.randn(B, c, w, w, generator=g)
    wt = torch.randn(c, c, 3, 3, generator=g) * 0.1
    ps, psh = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    gamma, beta = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g)
    for prologue in (False, True):
        xin = F.relu(x * ps.view(1, -1, 1, 1) + psh.view(1, -1, 1, 1)) if prologue else x
        ref = F.conv2d(xin, wt, None, 1, 1)
        out = torch.full((B, w, w, c), float("nan"), device="cuda")
        scratch = torch.zeros(int(lib.lc_conv_tc_scratch_floats(B, c, w)), device="cuda")
        stat = torch.zeros(4 * c, device="cuda")
        rc = lib.lc_conv3x3_tc(P(dev(nhwc(x))), P(dev(wt)), P(out), B, c, w, 0, P(dev(ps)) if prologue else None, P(dev(psh)) if prologue else None,
                               None, P(dev(gamma)), P(dev(beta)), None, P(stat), P(scratch), st())
        assert rc == 0
        torch.cuda.synchronize()
        assert int(scratch.view(torch.int32)[8]) == 0, "tensor-core barrier timed out"
        got = nchw(out).cpu()
        assert torch.isfinite(got).all()
        scale = ref.abs().max().item()
        err = (got - ref).abs().max().item()
        e_tr = (got - F.conv2d(tf32_trunc(xin), tf32_trunc(wt), None, 1, 1)).abs().max().item()
        e_rn = (got - F.conv2d(tf32_round(xin), tf32_round(wt), None, 1, 1)).abs().max().item()
        print(f"c={c} w={w} B={B} pro={prologue}: max|err| {err:.3e} (ref max {scale:.2f}); vs tf32-trunc operands {e_tr:.3e}; vs tf32-round {e_rn:.3e}")
        assert err <= 4e-3 * scale, (err, scale)
        mean, var = ref.mean((0, 2, 3)), ref.var((0, 2, 3), unbiased=False)
        close(stat[2 * c:3 * c], mean, 1e-2, 2e-3, "mean")
        close(stat[3 * c:], 1 / torch.sqrt(var + 1e-5), 1e-2, 1e-3, "invstd")


@pytest.mark.parametrize("c,w", [(16, 32), (32, 16), (64, 8)])
def test_conv3x3_tc_dgrad_with_addend(lib, c, w):
    B = 5
    g = torch.Generator().manual_seed(c)
    x = torch.randn(B, c, w, w, generator=g, requires_grad=True)
    wt = torch.randn(c, c, 3, 3, generator=g) * 0.1
    dy = torch.randn(B, c, w, w, generator=g)
    addend = torch.randn(B, c, w, w, generator=g)
    dx_ref, = torch.autograd.grad(F.conv2d(x, wt, None, 1, 1), [x], dy)
    dx = torch.empty(B, w, w, c, device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_tc_scratch_floats(B, c, w)), device="cuda")
    assert lib.lc_conv3x3_tc(P(dev(nhwc(dy))), P(dev(wt)), P(dx), B, c, w, 1, None, None, P(dev(nhwc(addend))), None, None, None, None, P(scratch), st()) == 0
    torch.cuda.synchronize()
    assert int(scratch.view(torch.int32)[8]) == 0
    ref = dx_ref + addend
    err = (nchw(dx).cpu() - ref).abs().max().item()
    assert err <= 4e-3 * dx_ref.abs().max().item(), err


@pytest.mark.parametrize("c,w", [(16, 32), (32, 16), (64, 8)])
@pytest.mark.parametrize("B", [2, 9])
def test_conv3x3_tc_wgrad(lib, c, w, B):
    g = torch.Generator().manual_seed(7 * c + B)
    x = torch.randn(B, c, w, w, generator=g)
    dy = torch.randn(B, c, w, w, generator=g)
    ps, psh = torch.rand(c, generator=g) + 0.5, torch.randn(c, generator=g) * 0.3
    for prologue in (False, True):
        xin = F.relu(x * ps.view(1, -1, 1, 1) + psh.view(1, -1, 1, 1)) if prologue else x
        ref = torch.nn.grad.conv2d_weight(xin, (c, c, 3, 3), dy, stride=1, padding=1)
        ref_tf = torch.nn.grad.conv2d_weight(xin.bfloat16().float(), (c, c, 3, 3), dy.bfloat16().float(), stride=1, padding=1)
        dw = torch.full((c, c, 3, 3), float("nan"), device="cuda")
        scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, c, c, w)), device="cuda")
        rc = lib.lc_conv3x3_wgrad_tc(P(dev(nhwc(x))), P(dev(nhwc(dy))), P(dw), B, c, w, P(dev(ps)) if prologue else None, P(dev(psh)) if prologue else None,
                                     P(scratch), st())
        assert rc == 0
        torch.cuda.synchronize()
        assert int(scratch.view(torch.int32)[8]) == 0, "tensor-core barrier timed out"
        got = dw.cpu()
        assert torch.isfinite(got).all()
        scale = ref.abs().max().item()
        err, err_tf = (got - ref).abs().max().item(), (got - ref_tf).abs().max().item()
        print(f"wgrad c={c} w={w} B={B} pro={prologue}: max|err| {err:.3e} (ref max {scale:.2f}); vs bf16-rounded operands {err_tf:.3e}")
        # BF16 operands (8 mantissa bits, round-to-nearest-even), fp32 accumulation: |err| <= 1e-2 * max|ref|; against an fp32
        # contraction of the same BF16-rounded operands only the summation order differs
        assert err <= 1e-2 * scale, (err, scale)
        assert err_tf <= 2e-4 * scale, (err_tf, scale)


@pytest.mark.parametrize("cin,wo", [(16, 16), (32, 8)])
@pytest.mark.parametrize("B", [1, 5, 37, 128])
def test_conv3x3s2_tc_forward(lib, cin, wo, B):
    """Stride-2 stage-transition conv on tcgen05 (parity-plane implicit GEMM, csrc/conv_s2_tc.cuh) against PyTorch fp32 conv2d(stride=2, padding=1)
    (resnet.py:341-343 at stride 2); same TF32 bound as the stride-1 kernel, BatchNorm statistics of the output included."""
    cout, win = 2 * cin, 2 * wo
    g = torch.Generator().manual_seed(1000 * cin + B)
    x = torch.randn(B, cin, win, win, generator=g)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    ref = F.conv2d(x, wt, None, 2, 1)
    out = torch.full((B, wo, wo, cout), float("nan"), device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, cin, cout, wo)), device="cuda")
    stat = torch.zeros(4 * cout, device="cuda")
    rc = lib.lc_conv3x3s2_tc(P(dev(nhwc(x))), P(dev(wt)), P(out), B, cin, wo, P(dev(gamma)), P(dev(beta)), None, P(stat), P(scratch), st())
    assert rc == 0
    torch.cuda.synchronize()
    assert int(scratch.view(torch.int32)[8]) == 0, "tensor-core barrier timed out"
    got = nchw(out).cpu()
    assert torch.isfinite(got).all()
    scale = ref.abs().max().item()
    err = (got - ref).abs().max().item()
    e_rn = (got - F.conv2d(tf32_round(x), tf32_round(wt), None, 2, 1)).abs().max().item()
    print(f"s2 cin={cin} wo={wo} B={B}: max|err| {err:.3e} (ref max {scale:.2f}); vs tf32-round operands {e_rn:.3e}")
    assert err <= 4e-3 * scale, (err, scale)
    assert e_rn <= 2e-5 * scale, (e_rn, scale)
    mean, var = ref.mean((0, 2, 3)), ref.var((0, 2, 3), unbiased=False)
    close(stat[2 * cout:3 * cout], mean, 1e-2, 2e-3, "mean")
    close(stat[3 * cout:], 1 / torch.sqrt(var + 1e-5), 1e-2, 1e-3, "invstd")


@pytest.mark.parametrize("cin,wo", [(16, 16), (32, 8)])
@pytest.mark.parametrize("B", [1, 5, 37, 128])
def test_conv3x3s2_tc_dgrad(lib, cin, wo, B):
    """Data gradient of the stride-2 conv on tcgen05 (nine tap groups into four parity-plane accumulators) against autograd of conv2d(stride=2)."""
    cout, win = 2 * cin, 2 * wo
    g = torch.Generator().manual_seed(77 * cin + B)
    x = torch.randn(B, cin, win, win, generator=g, requires_grad=True)
    wt = torch.randn(cout, cin, 3, 3, generator=g) * 0.1
    dy = torch.randn(B, cout, wo, wo, generator=g)
    dx_ref, = torch.autograd.grad(F.conv2d(x, wt, None, 2, 1), [x], dy)
    dx_rn, = torch.autograd.grad(F.conv2d(x, tf32_round(wt), None, 2, 1), [x], tf32_round(dy))
    dx = torch.full((B, win, win, cin), float("nan"), device="cuda")
    scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, cin, cout, wo)), device="cuda")
    assert lib.lc_conv3x3s2_dgrad_tc(P(dev(nhwc(dy))), P(dev(wt)), P(dx), B, cin, wo, P(scratch), st()) == 0
    torch.cuda.synchronize()
    assert int(scratch.view(torch.int32)[8]) == 0, "tensor-core barrier timed out"
    got = nchw(dx).cpu()
    assert torch.isfinite(got).all()
    scale = dx_ref.abs().max().item()
    err, e_rn = (got - dx_ref).abs().max().item(), (got - dx_rn).abs().max().item()
    print(f"s2 dgrad cin={cin} wo={wo} B={B}: max|err| {err:.3e} (ref max {scale:.2f}); vs tf32-round operands {e_rn:.3e}")
    assert err <= 4e-3 * scale, (err, scale)
    assert e_rn <= 2e-5 * scale, (e_rn, scale)
