"""Time the tcgen05 GEMM on the ViT-B/16 shapes (CUDA events, rotating buffers > L2) and report TFLOP/s vs the measured bf16 peak."""
import json, os, sys
sys.path.insert(0, '.')
import torch
from libcontinual_b200 import _lib
lib = _lib.load()
peak = 1652.1
try:
    peak = float(json.load(open('MEASURED_PEAKS.json'))['bf16_tflops'])
except Exception:
    pass
st = torch.cuda.current_stream().cuda_stream
for (M, N, K) in [(25216, 2304, 768), (25216, 768, 768), (25216, 3072, 768), (25216, 768, 3072), (28416, 2304, 768)]:
    nbuf = 4
    As = [torch.randn(M, K, device='cuda').bfloat16() for _ in range(nbuf)]
    Bs = [torch.randn(N, K, device='cuda').bfloat16() * 0.05 for _ in range(nbuf)]
    Cs = [torch.empty(M, N, device='cuda', dtype=torch.bfloat16) for _ in range(nbuf)]
    err = torch.zeros(4, dtype=torch.int32, device='cuda')
    def run(i):
        assert lib.lc_gemm_bf16(As[i].data_ptr(), K, 0, Bs[i].data_ptr(), K, 0, Cs[i].data_ptr(), N, 0, M, N, K, 1, None, None, 0, 0, None, 0, 1.0, err.data_ptr(), st) == 0
    for i in range(nbuf): run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for i in range(reps): run(i % nbuf)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
    # cuBLAS (torch.matmul) on the same shape, for context
    e0.record()
    for i in range(reps): torch.matmul(As[i % nbuf], Bs[i % nbuf].t())
    e1.record(); torch.cuda.synchronize()
    ms_t = e0.elapsed_time(e1) / reps
    print(f"M={M} N={N} K={K}: {ms*1e3:8.1f} us  {tf:7.1f} TFLOP/s  ({tf/peak*100:4.1f}% of measured bf16 peak {peak:.0f})   cuBLAS {2.0*M*N*K/(ms_t*1e-3)/1e12:7.1f} TFLOP/s  err={int(err[0])}")
