"""Per-tile role timeline of the persistent GEMM (needs the -DLC_GEMM_TIMING build: tools/_timing/liblc_timing.so).
   LC_B200_LIB=tools/_timing/liblc_timing.so python tools/gemm_timing.py M N K [bias] [gelu] [res]"""
import ctypes, os, sys
sys.path.insert(0, '.')
import numpy as np
import torch
from libcontinual_b200 import _lib
lib = _lib.load()
M, N, K = (int(v) for v in sys.argv[1:4])
flags = sys.argv[4:]
A = torch.randn(M, K, device='cuda').bfloat16(); B = (torch.randn(N, K, device='cuda') * 0.05).bfloat16()
C = torch.empty(M, N, device='cuda', dtype=torch.bfloat16); C2 = torch.empty_like(C)
bias = torch.randn(N, device='cuda')
res = torch.randn(M, N, device='cuda')
Cf = torch.empty(M, N, device='cuda')
err = torch.zeros(4, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
def run():
    if 'res' in flags:
        return lib.lc_gemm_bf16(A.data_ptr(), K, 0, B.data_ptr(), K, 0, Cf.data_ptr(), N, 0, M, N, K, 1, bias.data_ptr(), res.data_ptr(), N, 0, None, 1, 1.0, err.data_ptr(), st)
    return lib.lc_gemm_bf16(A.data_ptr(), K, 0, B.data_ptr(), K, 0, C.data_ptr(), N, 0, M, N, K, 1, bias.data_ptr() if 'bias' in flags else None, None, 0, 0,
                            C2.data_ptr() if 'gelu' in flags else None, 0, 1.0, err.data_ptr(), st)
for _ in range(3):
    assert run() == 0
torch.cuda.synchronize()
fn = ctypes.CDLL(os.environ["LC_B200_LIB"]).lc_debug_gemm_timing
buf = np.zeros(148 * 32 * 8, dtype=np.uint64)
assert fn(buf.ctypes.data_as(ctypes.c_void_p)) == 0
t = buf.reshape(148, 32, 8).astype(np.int64)
names = ["prod first", "prod last", "mma pre-wait", "mma start", "mma issued", "epi pre-wait", "epi start", "epi end"]
print(f"M={M} N={N} K={K} flags={flags}")
for cta in (0, 147):
    t0 = t[cta, 0, 0]
    print(f"CTA {cta} (ns since its first load issue)")
    for tl in range(32):
        if t[cta, tl, 7] == 0:
            break
        print(f"  tile {tl:2d}: " + "  ".join(f"{names[i]} {t[cta, tl, i] - t0:6d}" for i in range(8)))
valid = t[:, :, 7] > 0
epi = (t[:, :, 7] - t[:, :, 6])[valid]; mma = (t[:, :, 4] - t[:, :, 3])[valid]; wait = (t[:, :, 3] - t[:, :, 2])[valid]; ew = (t[:, :, 6] - t[:, :, 5])[valid]
print(f"tiles {valid.sum()}: epilogue {epi.mean():.0f} ns  mma issue span {mma.mean():.0f} ns  mma wait-for-accumulator {wait.mean():.0f} ns  epilogue wait-for-mma {ew.mean():.0f} ns")
