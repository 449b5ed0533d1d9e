"""Launch the stage-1 tensor-core conv a few times (for `ncu --set full -k regex:conv3x3_tc`)."""
import sys
sys.path.insert(0, '.')
import torch
from libcontinual_b200 import _lib
lib = _lib.load()
B, C, W = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = B * W * W * C
x = torch.randn(n, device='cuda'); y = torch.empty(n, device='cuda'); w = torch.randn(C, C, 3, 3, device='cuda') * 0.1
gamma = torch.ones(C, device='cuda'); beta = torch.zeros(C, device='cuda'); stat = torch.zeros(4 * C, device='cuda')
scratch = torch.zeros(int(lib.lc_conv_tc_scratch_floats(B, C, W)), device='cuda')
st = torch.cuda.current_stream().cuda_stream
for i in range(4):
    with_stats = i % 2 == 1
    rc = lib.lc_conv3x3_tc(x.data_ptr(), w.data_ptr(), y.data_ptr(), B, C, W, 0, None, None, None, gamma.data_ptr() if with_stats else None,
                           beta.data_ptr() if with_stats else None, None, stat.data_ptr() if with_stats else None, scratch.data_ptr(), st)
    assert rc == 0
torch.cuda.synchronize()
print('ok', int(scratch.view(torch.int32)[8]))
