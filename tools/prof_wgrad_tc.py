"""Launch the tensor-core weight gradient a few times (for `ncu --set full -k regex:wgrad3x3_tc`)."""
import sys
sys.path.insert(0, '.')
import torch
from libcontinual_b200 import _lib
lib = _lib.load()
B, C, W = 128, int(sys.argv[1]) if len(sys.argv) > 1 else 16, int(sys.argv[2]) if len(sys.argv) > 2 else 32
n = B * W * W * C
x = torch.randn(n, device='cuda'); dy = torch.randn(n, device='cuda'); dw = torch.empty(C * C * 9, device='cuda')
scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, C, C, W)) + C * C * 9 * 300, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for i in range(4):
    assert lib.lc_conv3x3_wgrad_tc(x.data_ptr(), dy.data_ptr(), dw.data_ptr(), B, C, W, None, None, scratch.data_ptr(), st) == 0
torch.cuda.synchronize()
print('ok', int(scratch.view(torch.int32)[8]))
