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
