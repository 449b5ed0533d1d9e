"""One GEMM launch per epilogue variant for ncu: python tools/prof_gemm.py M N K [bias] [gelu] [res]"""
import sys
sys.path.insert(0, '.')
import torch
from libcontinual_b200 import _lib
lib = _lib.load()
M, N, K = (int(v) for v in sys.argv[1:4])
flags = sys.argv[4:]
A = torch.randn(M, K, device='cuda').bfloat16(); B = (torch.randn(N, K, device='cuda') * 0.05).bfloat16()
C = torch.empty(M, N, device='cuda', dtype=torch.bfloat16); C2 = torch.empty_like(C)
bias = torch.randn(N, device='cuda'); res = torch.randn(M, N, device='cuda'); Cf = torch.empty(M, N, device='cuda')
err = torch.zeros(4, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    if 'res' in flags:
        rc = lib.lc_gemm_bf16(A.data_ptr(), K, 0, B.data_ptr(), K, 0, Cf.data_ptr(), N, 0, M, N, K, 1, bias.data_ptr(), res.data_ptr(), N, 0, None, 1, 1.0, err.data_ptr(), st)
    else:
        rc = lib.lc_gemm_bf16(A.data_ptr(), K, 0, B.data_ptr(), K, 0, C.data_ptr(), N, 0, M, N, K, 1, bias.data_ptr() if 'bias' in flags else None, None, 0, 0,
                              C2.data_ptr() if 'gelu' in flags else None, 0, 1.0, err.data_ptr(), st)
    assert rc == 0
torch.cuda.synchronize()
print("err", err.tolist())
