"""GPU: the generic layer kernels behind ResNet18 / AlexNet_TRGP (csrc/nn_ops.cuh, the implicit-GEMM convolution and split-K modes of the tcgen05 GEMM),
through the C ABI, against plain PyTorch fp32 / float64 CPU ops on the same seeded inputs.

Tolerances: BF16 operands with fp32 accumulation -> against a float64 contraction of the SAME BF16-rounded operands only the summation order differs
(<= 1e-4 of max|ref|, fp32 outputs); elementwise / reduction kernels in fp32: 1e-5 .. 1e-4; integer outputs (pool argmax) exact."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from tests.test_gpu_kernels import P, dev, lib, nchw, nhwc, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def _bf(t):
    return t.bfloat16().float()


CONV_CASES = [  # N, C, H, Cout, ks, stride, pad          (ResNet18 at 64x64: resnet.py:48-64,209-223)
    (4, 64, 32, 64, 3, 1, 1),       # layer1
    (4, 64, 32, 128, 3, 2, 1),      # layer2.0.conv1
    (4, 64, 32, 128, 1, 2, 0),      # layer2.0.downsample
    (6, 128, 16, 128, 3, 1, 1),     # layer2
    (6, 128, 16, 256, 3, 2, 1),     # layer3.0.conv1
    (5, 256, 8, 256, 3, 1, 1),      # layer3 (two images per 128-row tile, ragged batch)
    (9, 256, 8, 512, 3, 2, 1),      # layer4.0.conv1 (4x4 outputs: eight images per tile)
    (9, 512, 4, 512, 3, 1, 1),      # layer4
    (2, 64, 64, 64, 3, 1, 1),       # 64-wide rows (two rows per tile)
]


@pytest.mark.parametrize("N,C,H,Cout,ks,stride,pad", CONV_CASES)
def test_implicit_conv_gemm(lib, N, C, H, Cout, ks, stride, pad):
    from libcontinual_b200._lib import ConvDesc
    g = torch.Generator().manual_seed(N * 1000 + C + H + Cout)
    x = _bf(torch.randn(N, C, H, H, generator=g))
    w = _bf(torch.randn(Cout, C, ks, ks, generator=g) * (2.0 / (C * ks * ks)) ** 0.5)
    Ho = (H + 2 * pad - ks) // stride + 1
    ref = F.conv2d(x.double(), w.double(), None, stride, pad).float()
    xd = dev(nhwc(x).bfloat16())
    wk = torch.empty(Cout, ks * ks * C, dtype=torch.bfloat16, device="cuda")
    assert lib.lc_nn_pack_weight(P(dev(w)), Cout, C, ks, 0, 0, P(wk), ks * ks * C, st()) == 0
    res = torch.randn(N * Ho * Ho, Cout, generator=g)
    for with_res in (False, True):
        y = torch.full((N * Ho * Ho, Cout), float("nan"), device="cuda")
        err = torch.zeros(4, dtype=torch.int32, device="cuda")
        d = ConvDesc(X=P(xd), Wk=P(wk), Y=P(y), bias=None, residual=P(dev(res)) if with_res else None, ldc=Cout, ldr=Cout, N=N, H=H, W=H, C=C, Cout=Cout, ks=ks,
                     stride=stride, pad=pad, Ho=Ho, Wo=Ho, out_f32=1)
        assert lib.lc_conv_gemm_bf16(ctypes.byref(d), P(err), st()) == 0
        torch.cuda.synchronize()
        assert int(err[0]) == 0, "tensor-core barrier timed out"
        got = nchw(y.view(N, Ho, Ho, Cout)).cpu()
        want = ref + (nchw(res.view(N, Ho, Ho, Cout)) if with_res else 0)
        e = (got - want).abs().max().item()
        assert e <= 1e-4 * ref.abs().max().item() + 1e-5, (e, ref.abs().max().item())


def test_implicit_conv_dgrad_flipped_weights(lib):
    """Stride-1 data gradient = convolution of dY with the flipped, transposed filter (pack mode 2)."""
    from libcontinual_b200._lib import ConvDesc
    g = torch.Generator().manual_seed(3)
    N, C, H, Cout = 3, 128, 16, 64          # forward conv C -> Cout; the gradient conv runs Cout -> C
    x = torch.randn(N, C, H, H, generator=g, dtype=torch.float64, requires_grad=True)
    w = _bf(torch.randn(Cout, C, 3, 3, generator=g) * 0.05)
    dy = _bf(torch.randn(N, Cout, H, H, generator=g))
    ref, = torch.autograd.grad(F.conv2d(x, w.double(), None, 1, 1), [x], dy.double())
    wd = torch.empty(C, 9 * Cout, dtype=torch.bfloat16, device="cuda")
    assert lib.lc_nn_pack_weight(P(dev(w)), Cout, C, 3, 0, 2, P(wd), 9 * Cout, st()) == 0
    dx = torch.full((N * H * H, C), float("nan"), device="cuda")
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    d = ConvDesc(X=P(dev(nhwc(dy).bfloat16())), Wk=P(wd), Y=P(dx), bias=None, residual=None, ldc=C, ldr=C, N=N, H=H, W=H, C=Cout, Cout=C, ks=3, stride=1, pad=1,
                 Ho=H, Wo=H, out_f32=1)
    assert lib.lc_conv_gemm_bf16(ctypes.byref(d), P(err), st()) == 0
    torch.cuda.synchronize()
    assert int(err[0]) == 0
    got = nchw(dx.view(N, H, H, C)).cpu()
    assert (got - ref.float()).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-5


@pytest.mark.parametrize("M,N,K,ksplit", [(64, 576, 4096, 8), (128, 1152, 9216, 16), (512, 4608, 1024, 4), (64, 27, 8192, 13)])
def test_gemm_split_k(lib, M, N, K, ksplit):
    """Weight-gradient shape: short output, long contraction, split over ksplit fp32 partials."""
    from libcontinual_b200._lib import GemmDesc
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(M, K, generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, generator=g) * 0.5).bfloat16()
    ldc = (N + 3) // 4 * 4
    part = torch.full((ksplit, M, ldc), float("nan"), device="cuda")
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    d = GemmDesc(A=P(dev(A)), lda=K, B=P(dev(B)), ldb=K, C=P(part), ldc=ldc, M=M, N=N, K=K, batch_in=1, batch_out=1, out_f32=1, alpha=1.0, ksplit=ksplit,
                 strideC_split=M * ldc)
    assert lib.lc_gemm_bf16_ex(ctypes.byref(d), P(err), st()) == 0
    torch.cuda.synchronize()
    assert int(err[0]) == 0
    got = part[:, :, :N].sum(0).cpu()
    ref = (A.double() @ B.double().t()).float()
    assert (got - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-4


@pytest.mark.parametrize("N,C,H,ks,stride,pad,korder,kind", [(3, 3, 32, 4, 1, 0, 1, 1), (2, 64, 14, 3, 1, 0, 1, 0), (2, 128, 6, 2, 1, 0, 1, 0), (2, 3, 64, 3, 1, 1, 0, 1),
                                                             (2, 64, 16, 3, 2, 1, 0, 2)])
def test_im2col_and_transpose(lib, N, C, H, ks, stride, pad, korder, kind):
    g = torch.Generator().manual_seed(C + H)
    x = _bf(torch.randn(N, C, H, H, generator=g))
    Ho = (H + 2 * pad - ks) // stride + 1
    M, K = N * Ho * Ho, C * ks * ks
    Kp, ldT = (K + 7) // 8 * 8, (M + 7) // 8 * 8
    unf = F.unfold(x, ks, padding=pad, stride=stride)                     # [N][C*ks*ks][Ho*Wo], k = (c, kh, kw)
    ref = unf.permute(0, 2, 1).reshape(M, K)
    if korder == 0:
        ref = ref.view(M, C, ks * ks).permute(0, 2, 1).reshape(M, K)
    src = {0: lambda: nhwc(x).bfloat16(), 1: lambda: x.float(), 2: lambda: nhwc(x).float()}[kind]()
    col = torch.full((M, Kp), float("nan"), dtype=torch.bfloat16, device="cuda")
    colT = torch.full((Kp, ldT), float("nan"), dtype=torch.bfloat16, device="cuda")
    assert lib.lc_nn_im2col(P(dev(src)), kind, N, H, H, C, ks, stride, pad, korder, P(col), Kp, P(colT), ldT, Kp, st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(col[:, :K].float().cpu(), ref) and float(col[:, K:].float().abs().sum()) == 0.0
    assert torch.equal(colT[:K, :M].float().cpu(), ref.t()) and float(colT[K:].float().abs().sum()) == 0.0 and float(colT[:, M:].float().abs().sum()) == 0.0


@pytest.mark.parametrize("N,C,H,ks,stride,pad", [(3, 64, 32, 3, 1, 1), (2, 128, 16, 3, 2, 1), (5, 256, 8, 3, 1, 1), (7, 64, 16, 1, 2, 0), (3, 512, 4, 3, 1, 1)])
def test_im2col_transposed_fast_path(lib, N, C, H, ks, stride, pad):
    """The vectorised transposed patch matrix (NHWC BF16, C % 64 == 0, tap-major K): bit-identical to the generic kernel's definition."""
    g = torch.Generator().manual_seed(C * H + N)
    x = _bf(torch.randn(N, C, H, H, generator=g))
    Ho = (H + 2 * pad - ks) // stride + 1
    M, K = N * Ho * Ho, C * ks * ks
    ldT = (M + 7) // 8 * 8 + 8
    unf = F.unfold(x, ks, padding=pad, stride=stride)
    ref = unf.permute(0, 2, 1).reshape(M, C, ks * ks).permute(0, 2, 1).reshape(M, K)
    colT = torch.full((K, ldT), float("nan"), dtype=torch.bfloat16, device="cuda")
    assert lib.lc_nn_im2col(P(dev(nhwc(x).bfloat16())), 0, N, H, H, C, ks, stride, pad, 0, None, 0, P(colT), ldT, K, st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(colT[:, :M].float().cpu(), ref.t()) and float(colT[:, M:].float().abs().sum()) == 0.0


@pytest.mark.parametrize("N,C,H,ks,stride,pad,korder", [(2, 64, 14, 3, 1, 0, 1), (2, 64, 16, 3, 2, 1, 0), (3, 64, 16, 1, 2, 0, 0), (2, 128, 6, 2, 1, 0, 1)])
def test_col2im(lib, N, C, H, ks, stride, pad, korder):
    g = torch.Generator().manual_seed(7 * C + H)
    Ho = (H + 2 * pad - ks) // stride + 1
    M, K = N * Ho * Ho, C * ks * ks
    dcol = _bf(torch.randn(M, K, generator=g))
    addend = torch.randn(N, H, H, C, generator=g)
    d = dcol if korder == 1 else dcol.view(M, ks * ks, C).permute(0, 2, 1).reshape(M, K)       # -> (c, kh, kw) for F.fold
    ref = F.fold(d.view(N, Ho * Ho, K).permute(0, 2, 1).double(), (H, H), ks, padding=pad, stride=stride).float()
    dx = torch.full((N, H, H, C), float("nan"), device="cuda")
    assert lib.lc_nn_col2im(P(dev(dcol.bfloat16())), K, P(dev(addend)), P(dx), N, H, H, C, ks, stride, pad, korder, st()) == 0
    torch.cuda.synchronize()
    assert torch.allclose(nchw(dx).cpu(), ref + nchw(addend), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("M,C", [(4 * 29 * 29, 64), (64, 2048), (8 * 32 * 32, 128), (300, 512)])
def test_batchnorm_forward_backward(lib, M, C):
    """nn.BatchNorm (train) + ReLU on a [M][C] matrix: statistics, running update, output, and the full backward vs autograd in float64."""
    g = torch.Generator().manual_seed(M + C)
    y = torch.randn(M, C, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.2
    res = torch.randn(M, C, generator=g)
    gout = torch.randn(M, C, generator=g)
    yd = y.double().requires_grad_(True); gd = gamma.double().requires_grad_(True); bd = beta.double().requires_grad_(True)
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    out_ref = F.relu(F.batch_norm(yd, rm, rv, gd, bd, True, 0.1, 1e-5) + res.double())
    dy_ref, dg_ref, db_ref = torch.autograd.grad(out_ref, [yd, gd, bd], gout.double())
    scratch = torch.zeros(int(lib.lc_nn_bn_scratch_floats(C)), device="cuda")
    aff = torch.zeros(4 * C, device="cuda")
    running = torch.cat([torch.zeros(C), torch.ones(C)]).cuda()
    ydv = dev(y)
    assert lib.lc_nn_bn_stats(P(ydv), M, C, P(dev(gamma)), P(dev(beta)), 1e-5, 0.1, P(running), P(aff), P(scratch), st()) == 0
    out = torch.empty(M, C, device="cuda"); outb = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    assert lib.lc_nn_bn_act(P(ydv), P(aff), P(dev(res)), None, M, C, 1, 0.0, None, 0, P(outb), P(out), st()) == 0
    torch.cuda.synchronize()
    assert torch.allclose(running[:C].cpu().double(), rm, rtol=1e-4, atol=1e-5) and torch.allclose(running[C:].cpu().double(), rv, rtol=1e-4, atol=1e-5)
    assert torch.allclose(out.cpu(), out_ref.float(), rtol=1e-4, atol=1e-4)
    assert torch.allclose(outb.float().cpu(), out_ref.float(), rtol=1e-2, atol=1e-2)
    dyb = torch.empty(M, C, device="cuda", dtype=torch.bfloat16); dyf = torch.empty(M, C, device="cuda"); dz = torch.empty(M, C, device="cuda")
    dgam, dbet = torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    assert lib.lc_nn_bn_backward(P(dev(gout)), P(out), None, 1.0, P(ydv), P(aff), M, C, P(dgam), P(dbet), P(dyb), P(dyf), P(dz), P(scratch), st()) == 0
    torch.cuda.synchronize()
    scale = dy_ref.abs().max().item()
    assert (dyf.cpu() - dy_ref.float()).abs().max().item() <= 2e-4 * scale
    assert torch.allclose(dgam.cpu(), dg_ref.float(), rtol=1e-3, atol=1e-3 * dg_ref.abs().max().item())
    assert torch.allclose(dbet.cpu(), db_ref.float(), rtol=1e-3, atol=1e-3 * db_ref.abs().max().item())
    assert torch.equal(dz.cpu(), torch.where(out_ref > 0, gout.double(), torch.zeros((), dtype=torch.float64)).float())


def test_dropout_mask_is_reproducible_and_unbiased(lib):
    M, C, p = 4096, 256, 0.5
    rng = torch.tensor([12345, 7], dtype=torch.int64, device="cuda")
    keep = torch.empty(M * C, dtype=torch.uint8, device="cuda")
    assert lib.lc_nn_dropout_mask(P(rng), 3, p, M * C, P(keep), st()) == 0
    y = torch.rand(M, C, device="cuda") + 0.1
    aff = torch.cat([torch.ones(C), torch.zeros(3 * C)]).cuda()
    out = torch.empty(M, C, device="cuda")
    assert lib.lc_nn_bn_act(P(y), P(aff), None, None, M, C, 1, p, P(rng), 3, None, P(out), st()) == 0
    torch.cuda.synchronize()
    k = keep.view(M, C).bool()
    assert torch.equal(out > 0, k) and torch.allclose(out[k], y[k] * 2.0)
    assert abs(float(k.float().mean()) - 0.5) < 5e-3
    keep2 = torch.empty_like(keep)
    assert lib.lc_nn_rng_advance(P(rng), st()) == 0 and lib.lc_nn_dropout_mask(P(rng), 3, p, M * C, P(keep2), st()) == 0
    torch.cuda.synchronize()
    assert int(rng[1]) == 8 and 0.45 < float((keep2 != keep).float().mean()) < 0.55       # a new step draws an independent mask


@pytest.mark.parametrize("N,C,H,k,stride,pad", [(3, 64, 29, 2, 2, 0), (2, 64, 64, 3, 2, 1), (2, 256, 5, 2, 2, 0)])
def test_maxpool_forward_backward(lib, N, C, H, k, stride, pad):
    g = torch.Generator().manual_seed(H)
    x = torch.randn(N, C, H, H, generator=g, requires_grad=True)
    ref = F.max_pool2d(x, k, stride, pad)
    Ho = ref.shape[2]
    gout = torch.randn(N, C, Ho, Ho, generator=g)
    dref, = torch.autograd.grad(ref, [x], gout)
    out = torch.empty(N, Ho, Ho, C, device="cuda"); idx = torch.empty(N, Ho, Ho, C, dtype=torch.uint8, device="cuda")
    xin = dev(nhwc(x.detach()))
    assert lib.lc_nn_maxpool_forward(P(xin), N, H, H, C, k, stride, pad, P(out), None, P(idx), st()) == 0
    dx = torch.empty(N, H, H, C, device="cuda")
    assert lib.lc_nn_maxpool_backward(P(dev(nhwc(gout))), P(idx), N, H, H, C, k, stride, pad, P(dx), st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(nchw(out).cpu(), ref.detach()) and torch.allclose(nchw(dx).cpu(), dref, rtol=1e-6, atol=1e-6)


def test_pack_weight_wgrad_reduce_cast_transpose(lib):
    g = torch.Generator().manual_seed(1)
    Cout, Cin, ks = 24, 16, 3
    w = torch.randn(Cout, Cin, ks, ks, generator=g)
    K = Cin * ks * ks
    wd = dev(w)
    for korder in (0, 1):
        out = torch.empty(Cout, K + 8, dtype=torch.bfloat16, device="cuda")
        assert lib.lc_nn_pack_weight(P(wd), Cout, Cin, ks, korder, 0, P(out), K + 8, st()) == 0
        ref = w.reshape(Cout, K) if korder == 1 else w.permute(0, 2, 3, 1).reshape(Cout, K)
        outT = torch.empty(K, Cout, dtype=torch.bfloat16, device="cuda")
        assert lib.lc_nn_pack_weight(P(wd), Cout, Cin, ks, korder, 1, P(outT), Cout, st()) == 0
        torch.cuda.synchronize()
        assert torch.equal(out[:, :K].float().cpu(), _bf(ref)) and float(out[:, K:].float().abs().sum()) == 0 and torch.equal(outT.float().cpu(), _bf(ref).t())
        part = torch.randn(5, Cout, K + 4, generator=g)
        dw = torch.empty(Cout, Cin, ks, ks, device="cuda")
        assert lib.lc_nn_wgrad_reduce(P(dev(part)), 5, Cout, Cin, ks, korder, K + 4, P(dw), st()) == 0
        torch.cuda.synchronize()
        s = part[:, :, :K].sum(0)
        want = s.view(Cout, Cin, ks, ks) if korder == 1 else s.view(Cout, ks, ks, Cin).permute(0, 3, 1, 2)
        assert torch.allclose(dw.cpu(), want, rtol=1e-6, atol=1e-6)
    src = torch.randn(70, 50, generator=g)
    a = torch.empty(70, 56, dtype=torch.bfloat16, device="cuda"); b = torch.empty(50, 72, dtype=torch.bfloat16, device="cuda")
    assert lib.lc_nn_cast_transpose(P(dev(src)), 70, 50, P(a), 56, P(b), 72, st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(a[:, :50].float().cpu(), _bf(src)) and torch.equal(b[:, :70].float().cpu(), _bf(src).t())
    assert float(a[:, 50:].float().abs().sum()) == 0 and float(b[:, 70:].float().abs().sum()) == 0
