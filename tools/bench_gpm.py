"""Times the GPM projection of the five AlexNet_TRGP layers (one training step's worth: gpm.py:78-81), CUDA-core kernel vs tensor-core split kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from libcontinual_b200 import _lib
from libcontinual_b200.gpm import GPMProjector
lib = _lib.load()
shapes = [(64, 48), (128, 576), (256, 512), (2048, 1024), (2048, 2048)]
feats = [torch.linalg.qr(torch.randn(D, max(4, D // 10), device="cuda"))[0] for _, D in shapes]
proj = GPMProjector(feats)
Ms = [(f @ f.T).contiguous() for f in feats]
gs = [torch.randn(r, D, device="cuda") for r, D in shapes]
st = torch.cuda.current_stream().cuda_stream
def run_simt():
    for g, M, (r, D) in zip(gs, Ms, shapes):
        assert lib.lc_gpm_project(g.data_ptr(), M.data_ptr(), r, D, st) == 0
def run_tc():
    for i, g in enumerate(gs):
        proj.project_(i, g)
for name, fn in (("cuda-core fp32 (lc_gpm_project)", run_simt), ("tcgen05 bf16x3 split (lc_gpm_project_tc)", run_tc)):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flop = sum(2.0 * r * D * D for r, D in shapes)
    print(f"{name}: {ms * 1e3:.0f} us per step (5 layers, {flop / 1e9:.1f} GFLOP dense) -> {flop / ms / 1e9:.1f} TFLOP/s effective")
