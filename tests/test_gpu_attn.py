"""Fused tcgen05 attention (forward and backward) against a plain PyTorch fp32 evaluation of the same op on BF16-rounded inputs
(core/model/backbone/transformer.py:169-197).  Tolerance: BF16 probabilities / outputs -> 1e-2 relative L2."""
import pytest
import torch

from libcontinual_b200 import _lib

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def ref_attention(qkv, B, T, H):
    q, k, v = qkv.float().reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    p = ((q @ k.transpose(-2, -1)) * 0.125).softmax(-1)
    return (p @ v).transpose(1, 2).reshape(B, T, H * 64), p


@pytest.mark.parametrize("B,T,H", [(2, 197, 12), (3, 222, 12), (1, 128, 2), (2, 77, 3), (1, 256, 1), (1, 130, 4)])
def test_attention_forward_backward(B, T, H):
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1000 + T)
    qkv = (torch.randn(B, T, 3, H, 64, generator=g) * 1.5).to(dev).bfloat16().contiguous()
    dout = torch.randn(B, T, H * 64, generator=g).to(dev).bfloat16().contiguous()
    out = torch.full((B, T, H * 64), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.lc_attn_forward(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, err.data_ptr(), st), "attn_forward")
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    x = qkv.float().requires_grad_(True)
    ref, p = ref_attention(x, B, T, H)
    e = rel_l2(out.float(), ref)
    print(f"B{B} T{T} H{H}: forward rel-L2 {e:.2e}")
    assert e < 1e-2
    # log-sum-exp (base 2) of the scaled scores
    q, k, _ = x.detach().reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    lse_ref = torch.logsumexp((q @ k.transpose(-2, -1)) * 0.125, dim=-1) * 1.4426950408889634
    assert float((lse - lse_ref).abs().max()) < 2e-2
    ref.backward(dout.float())
    dqkv = torch.full((B, T, 3, H, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    rowdot = torch.zeros(B, H, T, device=dev)
    _lib.check(lib.lc_attn_backward(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), rowdot.data_ptr(), dqkv.data_ptr(), B, T, H,
                                    err.data_ptr(), st), "attn_backward")
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    gref = x.grad.reshape(B, T, 3, H, 64)
    for i, nm in enumerate(("dQ", "dK", "dV")):
        e = rel_l2(dqkv[:, :, i].float(), gref[:, :, i])
        print(f"   {nm} rel-L2 {e:.2e}")
        assert e < 1.5e-2, (nm, e)


@pytest.mark.parametrize("B,T,H,P", [(2, 197, 12, 10), (3, 197, 12, 3), (2, 197, 12, 4), (1, 50, 2, 16)])
def test_attention_prefix_kv_forward_backward(B, T, H, P):
    """`MultiHeadAttention.forward(prompt=(pk, pv))` (transformer.py:175-180): P prefix keys / values per image in front of the token keys."""
    lib = _lib.load()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(2000 + T + P)
    qkv = (torch.randn(B, T, 3, H, 64, generator=g) * 1.5).to(dev).bfloat16().contiguous()
    pk = (torch.randn(B, P, H * 64, generator=g) * 1.5).to(dev).bfloat16().contiguous()
    pv = (torch.randn(B, P, H * 64, generator=g) * 1.5).to(dev).bfloat16().contiguous()
    dout = torch.randn(B, T, H * 64, generator=g).to(dev).bfloat16().contiguous()
    out = torch.full((B, T, H * 64), float("nan"), device=dev, dtype=torch.bfloat16)
    lse = torch.zeros(B, H, T, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(lib.lc_attn_forward_prefix(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, pk.data_ptr(), pv.data_ptr(), P, err.data_ptr(), st), "fwd")
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    x = qkv.float().requires_grad_(True)
    xk = pk.float().requires_grad_(True); xv = pv.float().requires_grad_(True)
    q, k, v = x.reshape(B, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    k = torch.cat((xk.reshape(B, P, H, 64).permute(0, 2, 1, 3), k), dim=2)
    v = torch.cat((xv.reshape(B, P, H, 64).permute(0, 2, 1, 3), v), dim=2)
    ref = (((q @ k.transpose(-2, -1)) * 0.125).softmax(-1) @ v).transpose(1, 2).reshape(B, T, H * 64)
    e = rel_l2(out.float(), ref)
    print(f"B{B} T{T} H{H} P{P}: forward rel-L2 {e:.2e}")
    assert e < 1e-2
    ref.backward(dout.float())
    dqkv = torch.full((B, T, 3, H, 64), float("nan"), device=dev, dtype=torch.bfloat16)
    dpk = torch.full((B, P, H * 64), float("nan"), device=dev); dpv = torch.full((B, P, H * 64), float("nan"), device=dev)
    _lib.check(lib.lc_attn_backward_prefix(qkv.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, T, H, pk.data_ptr(), pv.data_ptr(),
                                           dpk.data_ptr(), dpv.data_ptr(), P, err.data_ptr(), st), "bwd")
    torch.cuda.synchronize()
    assert int(err.item()) == 0
    gref = x.grad.reshape(B, T, 3, H, 64)
    for i, nm in enumerate(("dQ", "dK", "dV")):
        e = rel_l2(dqkv[:, :, i].float(), gref[:, :, i])
        print(f"   {nm} rel-L2 {e:.2e}")
        assert e < 1.5e-2, (nm, e)
    for nm, got, want in (("dpk", dpk, xk.grad), ("dpv", dpv, xv.grad)):
        e = rel_l2(got, want)
        print(f"   {nm} rel-L2 {e:.2e}")
        assert e < 1.5e-2, (nm, e)
