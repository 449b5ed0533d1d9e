"""GPU: tcgen05 BF16 GEMM (TMA + 128B swizzle + TMEM) through the C ABI vs a float64 matmul of the same BF16-rounded operands.
Products of bf16 values are exact in fp32, so only the fp32 accumulation order differs: |err| <= 1e-5 * K * max|a||b|-ish; asserted
as rel 2e-3 of max|ref| for bf16 outputs (output rounding 2^-9) and 1e-4 for fp32 outputs."""
import pytest
import torch

from tests.test_gpu_kernels import P, dev, lib, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def run_gemm(lib, A, B, *, bias=None, residual=None, gelu=False, out_f32=False, alpha=1.0, batch=1):
    """A [batch, M, K], B [batch, N, K] bf16 (CPU) -> C [batch, M, N] (stored with the row stride padded to a multiple of 16)."""
    b, M, K = A.shape
    N = B.shape[1]
    ld = (N + 15) // 16 * 16
    Ad, Bd = dev(A), dev(B)
    C = torch.full((b, M, ld), float("nan"), device="cuda", dtype=torch.float32 if out_f32 else torch.bfloat16)
    C2 = torch.full((b, M, ld), float("nan"), device="cuda", dtype=torch.bfloat16) if gelu else None
    err = torch.zeros(4, dtype=torch.int32, device="cuda")
    rc = lib.lc_gemm_bf16(P(Ad), K, M * K, P(Bd), K, N * K, P(C), ld, M * ld, M, N, K, b, P(dev(bias)) if bias is not None else None,
                          P(dev(residual)) if residual is not None else None, N, M * N, P(C2), int(out_f32), float(alpha), P(err), st())
    assert rc == 0
    torch.cuda.synchronize()
    assert int(err[0]) == 0, "tensor-core barrier timed out"
    return C[:, :, :N].float().cpu(), (C2[:, :, :N].float().cpu() if gelu else None)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 256, 768), (394, 2304, 768), (444, 768, 3072), (197, 768, 768), (130, 128, 128)])
def test_gemm_plain(lib, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(1, M, K, generator=g) * 0.5).bfloat16()
    B = (torch.randn(1, N, K, generator=g) * 0.05).bfloat16()
    C, _ = run_gemm(lib, A, B)
    ref = (A.double() @ B.double().transpose(1, 2)).float()
    err = (C - ref).abs().max().item()
    assert err <= 8e-3 * ref.abs().max().item(), (err, ref.abs().max().item())
    Cf, _ = run_gemm(lib, A, B, out_f32=True)
    errf = (Cf - ref).abs().max().item()
    assert errf <= 1e-4 * ref.abs().max().item() + 1e-5, (errf, ref.abs().max().item())


def test_gemm_epilogues(lib):
    g = torch.Generator().manual_seed(5)
    M, N, K = 394, 768, 768
    A = (torch.randn(1, M, K, generator=g) * 0.5).bfloat16()
    B = (torch.randn(1, N, K, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, generator=g)
    res = torch.randn(1, M, N, generator=g)
    ref = (A.double() @ B.double().transpose(1, 2)).float() + bias
    C, _ = run_gemm(lib, A, B, bias=bias, residual=res, out_f32=True)
    assert (C - (ref + res)).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-4
    Z, Gz = run_gemm(lib, A, B, bias=bias, gelu=True)
    assert (Z - ref).abs().max().item() <= 8e-3 * ref.abs().max().item()
    assert (Gz - torch.nn.functional.gelu(ref)).abs().max().item() <= 8e-3 * ref.abs().max().item()


def test_gemm_batched_ragged_and_scaled(lib):
    """Attention-shaped: per (batch*head) Q K^T with 222 keys (N not a multiple of 16), K = 64, alpha = 1/8."""
    g = torch.Generator().manual_seed(9)
    b, T, d = 6, 222, 64
    Q = (torch.randn(b, T, d, generator=g)).bfloat16()
    Kk = (torch.randn(b, T, d, generator=g)).bfloat16()
    S, _ = run_gemm(lib, Q, Kk, out_f32=True, alpha=0.125)
    ref = (Q.double() @ Kk.double().transpose(1, 2)).float() * 0.125
    assert torch.isfinite(S).all()
    assert (S - ref).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-5


def test_gemm_rejects_misaligned_rows(lib):
    A = torch.zeros(1, 128, 64, dtype=torch.bfloat16, device="cuda"); B = torch.zeros(1, 222, 64, dtype=torch.bfloat16, device="cuda")
    C = torch.zeros(1, 128, 222, device="cuda")
    assert lib.lc_gemm_bf16(P(A), 64, 0, P(B), 64, 0, P(C), 222, 0, 128, 222, 64, 1, None, None, 0, 0, None, 1, 1.0, None, st()) == -22


def test_gemm_gelu_derivative_side_tensor(lib):
    """gelu_mode bit 0: the out2 variant stores GELU'(pre-activation) in C (and GELU in out2); the backward variant multiplies by it (what the
    frozen-backbone fc2 data gradient needs).  bit 1: C is not written at all."""
    import ctypes
    from libcontinual_b200._lib import GemmDesc
    g = torch.Generator().manual_seed(15)
    M, N, K = 394, 768, 256
    A = (torch.randn(M, K, generator=g) * 0.5).bfloat16(); B = (torch.randn(N, K, generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, generator=g) * 0.1
    Ad, Bd, bd = dev(A), dev(B), dev(bias)
    C = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16); C2 = torch.full((M, N), 7.0, device="cuda", dtype=torch.bfloat16)
    err = torch.zeros(4, dtype=torch.int32, device="cuda")

    def call(mode, C_, C2_, aux=None, Aop=Ad, Kk=K):
        d = GemmDesc()
        d.A, d.lda, d.B, d.ldb, d.C, d.ldc = P(Aop), Kk, P(Bd), K, P(C_), N
        d.bias = P(bd) if aux is None else None
        d.out2 = None if C2_ is None else P(C2_)
        d.gelu_bwd_aux = None if aux is None else P(aux)
        d.M, d.N, d.K, d.batch_in, d.batch_out, d.out_f32, d.alpha, d.gelu_mode = M, N, K, 1, 1, 0, 1.0, mode
        assert lib.lc_gemm_bf16_ex(ctypes.byref(d), P(err), st()) == 0
        torch.cuda.synchronize()
        assert int(err[0]) == 0

    x = (A.double() @ B.double().T + bias.double()).requires_grad_(True)
    y = torch.nn.functional.gelu(x)
    y.sum().backward()
    call(1, C, C2)
    assert (C2.float().cpu() - y.detach().float()).abs().max().item() <= 8e-3 * y.abs().max().item()
    assert (C.float().cpu() - x.grad.float()).abs().max().item() <= 8e-3                       # GELU' in [-0.13, 1.13], BF16 rounding 2^-9
    call(2, C, C2)                                                                              # C untouched
    assert torch.equal(C.cpu().float(), C.cpu().float()) and (C2.float().cpu() - y.detach().float()).abs().max().item() <= 8e-3 * y.abs().max().item()
    keep = C.clone()
    C.fill_(3.0); call(2, C, C2)
    assert float((C.float() - 3.0).abs().max()) == 0.0
    # backward: D = (A B^T) * aux  with aux = GELU' (mode 1) equals D = (A B^T) * GELU'(pre) with aux = pre (mode 0)
    pre = x.detach().float().bfloat16()
    D1 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16); D0 = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    call(1, D1, None, aux=keep)
    call(0, D0, None, aux=dev(pre))
    ref = (A.double() @ B.double().T) * x.grad
    assert (D1.float().cpu() - ref.float()).abs().max().item() <= 1.5e-2 * ref.abs().max().item()
    assert (D0.float().cpu() - ref.float()).abs().max().item() <= 1.5e-2 * ref.abs().max().item()


@pytest.mark.parametrize("M,N,K", [(4864, 1024, 256), (4997, 1000, 192), (2560, 2304, 768)])
def test_gemm_cta_pair_variant(lib, M, N, K, monkeypatch):
    """Shapes with >= 74 blocks of 256 x 256 run the cta_group::2 kernel (2-CTA clusters, each CTA stages its own 128 rows of A and half of the B panel):
    same contract, checked against float64 for plain / bias + residual fp32 / bias + GELU outputs, a ragged M (last pair: rows past M in both halves) and
    a ragged N, and against the single-CTA kernel (LC_GEMM_CG2 is read once per process, so the comparison is against float64 only)."""
    assert ((M + 255) // 256) * ((N + 255) // 256) >= 74
    g = torch.Generator().manual_seed(M + N + K)
    A = (torch.randn(1, M, K, generator=g) * 0.5).bfloat16()
    B = (torch.randn(1, N, K, generator=g) * 0.05).bfloat16()
    bias = torch.randn(N, generator=g)
    res = torch.randn(1, M, N, generator=g)
    ref = (A.double() @ B.double().transpose(1, 2)).float()
    C, _ = run_gemm(lib, A, B)
    assert (C - ref).abs().max().item() <= 8e-3 * ref.abs().max().item()
    Cf, _ = run_gemm(lib, A, B, bias=bias, residual=res, out_f32=True)
    assert (Cf - (ref + bias + res)).abs().max().item() <= 1e-4 * ref.abs().max().item() + 1e-4
    Z, Gz = run_gemm(lib, A, B, bias=bias, gelu=True)
    assert (Z - (ref + bias)).abs().max().item() <= 8e-3 * (ref + bias).abs().max().item()
    assert (Gz - torch.nn.functional.gelu(ref + bias)).abs().max().item() <= 8e-3 * (ref + bias).abs().max().item()
