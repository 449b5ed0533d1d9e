import sys; sys.path.insert(0, '.')
import torch, torch.nn.functional as F
from libcontinual_b200 import _lib
lib = _lib.load()
c, w, B = int(sys.argv[1]), int(sys.argv[2]), 2
g = torch.Generator().manual_seed(1)
x = torch.randn(B, c, w, w, generator=g); dy = torch.randn(B, c, w, w, generator=g)
ref = torch.nn.grad.conv2d_weight(x, (c, c, 3, 3), dy, stride=1, padding=1)
xd = x.permute(0,2,3,1).contiguous().cuda(); dyd = dy.permute(0,2,3,1).contiguous().cuda()
dw = torch.full((c, c, 3, 3), float('nan'), device='cuda')
scratch = torch.zeros(int(lib.lc_conv_scratch_floats(B, c, c, w)), device='cuda')
rc = lib.lc_conv3x3_wgrad_tc(xd.data_ptr(), dyd.data_ptr(), dw.data_ptr(), B, c, w, None, None, scratch.data_ptr(), torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
got = dw.cpu()
print('rc', rc, 'err flag', int(scratch.view(torch.int32)[8]), 'got norm', float(got.norm()), 'ref norm', float(ref.norm()), 'zeros frac', float((got == 0).float().mean()))
def corr(a, b): return float((a * b).sum() / (a.norm() * b.norm() + 1e-30))
print('corr ref', corr(got, ref))
print('corr ref^T(co<->ci)', corr(got, ref.transpose(0, 1)))
print('corr tap-flipped', corr(got, ref.flip(2, 3)))
print('corr T + flip', corr(got, ref.transpose(0, 1).flip(2, 3)))
for t in range(9):
    print('tap', t, 'corr', corr(got[:, :, t // 3, t % 3], ref[:, :, t // 3, t % 3]), 'norm ratio', float(got[:, :, t // 3, t % 3].norm() / ref[:, :, t // 3, t % 3].norm()))
print(got[0, :4, 1, 1], ref[0, :4, 1, 1])
print(got[:4, 0, 1, 1], ref[:4, 0, 1, 1])
