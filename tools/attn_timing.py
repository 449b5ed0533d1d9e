"""Phase timeline of the fused attention kernels (needs the -DLC_ATTN_TIMING build: tools/_timing/liblc_attn_timing.so, made by tools/build_timing.sh).
   LC_B200_LIB=tools/_timing/liblc_attn_timing.so python tools/attn_timing.py [B] [T]"""
import ctypes, os, sys
sys.path.insert(0, '.')
import numpy as np
import torch
from libcontinual_b200 import _lib
lib = _lib.load()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
T = int(sys.argv[2]) if len(sys.argv) > 2 else 197
H = 12
qkv = (torch.randn(B, T, 3, H, 64, device='cuda') * 1.0).bfloat16()
dout = torch.randn(B, T, H * 64, device='cuda').bfloat16()
out = torch.empty(B, T, H * 64, device='cuda', dtype=torch.bfloat16)
dqkv = torch.empty_like(qkv)
lse = torch.zeros(B, H, T, device='cuda'); rowdot = torch.zeros(B, H, T, device='cuda')
err = torch.zeros(1, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
dbg = ctypes.CDLL(os.environ["LC_B200_LIB"]).lc_debug_attn_timing
buf = np.zeros(2048 * 16, dtype=np.uint64)

def report(names, what):
    assert dbg(buf.ctypes.data_as(ctypes.c_void_p), 0) == 0
    t = buf.reshape(2048, 16).astype(np.int64)
    valid = t[:, 0] > 0
    t = t[valid]
    last = max(i for i in range(16) if (t[:, i] > 0).any())
    print(f"{what}: {valid.sum()} CTAs sampled; kernel span {(t[:, last].max() - t[:, 0].min()) / 1e3:.1f} us; per-CTA total median {np.median(t[:, last] - t[:, 0]) / 1e3:.2f} us")
    prev = 0
    for i in range(1, 16):
        if not (t[:, i] > 0).all():
            continue
        d = (t[:, i] - t[:, prev]) / 1e3
        print(f"   {names.get(i, str(i)):34s} median {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f} us")
        prev = i

for _ in range(2):
    assert lib.lc_attn_forward(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, err.data_ptr(), st) == 0
torch.cuda.synchronize()
assert dbg(None, 1) == 0
assert lib.lc_attn_forward(qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, err.data_ptr(), st) == 0
torch.cuda.synchronize()
report({1: "stage Q,K,V + sync", 2: "S = Q K^T", 3: "row max", 4: "exp + P tile + sync", 5: "O = P V", 6: "normalise + store"}, f"forward B{B} T{T}")
for _ in range(2):
    assert lib.lc_attn_backward(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), rowdot.data_ptr(), dqkv.data_ptr(), B, T, H, err.data_ptr(), st) == 0
torch.cuda.synchronize()
assert dbg(None, 1) == 0
assert lib.lc_attn_backward(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), rowdot.data_ptr(), dqkv.data_ptr(), B, T, H, err.data_ptr(), st) == 0
torch.cuda.synchronize()
nm = {1: "stage K,V,Q,dO + sync", 2: "S = Q K^T", 3: "P tile + sync", 4: "dP = dO V^T", 5: "D, dS tile + sync", 6: "dQ, dK, dV MMAs", 7: "dQ store + sync",
      8: "q1: stage Q,dO + sync", 9: "q1: S", 10: "q1: P tile", 11: "q1: dP", 12: "q1: D, dS", 13: "q1: MMAs", 14: "q1: dQ store", 15: "dK, dV store"}
report(nm, f"backward B{B} T{T}")
print("err", int(err.item()))
