"""Debug: per-layer forward (and optionally backward) comparison of ResNet18Engine with the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F
from oracle import port
from tests.golden_util import synth_batch, synth_resnet18_state
from tests.test_gpu_resnet18 import make_backbone, rel_l2

p, b, _, _ = synth_resnet18_state(1818, 20)
B = 8
bb = make_backbone(p, b, max_batch=B)
eng = bb.engine
x, _ = synth_batch(77, B, 0, 10, img=64)
bb.train()
with torch.no_grad():
    out = bb(x.cuda())
ws = eng.ws
nchw = lambda t, n, h, c: t[:n * h * h].view(n, h, h, c).permute(0, 3, 1, 2).float().cpu()
# oracle step by step
ob = {k: v.clone() for k, v in b.items()}
h = F.conv2d(x, p["conv1.0.weight"], None, 1, 1)
print("stem conv y0", rel_l2(nchw(ws.y["conv1.0"], B, 64, 64), h))
h = F.relu(port._bn(h, p, ob, "conv1.1", True))
print("stem act a0", rel_l2(nchw(ws.a0, B, 64, 64), h))
h = F.max_pool2d(h, 3, 2, 1)
print("pool x0", rel_l2(nchw(ws.x0_f32, B, 32, 64), h), rel_l2(nchw(ws.x0_bf, B, 32, 64), h))
for li, stride in enumerate((1, 2, 2, 2), start=1):
    for k in range(2):
        pre = f"layer{li}.{k}"
        s = stride if k == 0 else 1
        identity = h
        y = F.conv2d(h, p[pre + ".conv1.weight"], None, s, 1)
        c1 = eng.conv_by_name[pre + ".conv1"]
        print(pre, "conv1 y", rel_l2(nchw(ws.y[pre + ".conv1"], B, c1.Ho, c1.cout), y))
        y = F.relu(port._bn(y, p, ob, pre + ".bn1", True))
        print(pre, "a1", rel_l2(nchw(ws.a1[pre], B, c1.Ho, c1.cout), y))
        y = F.conv2d(y, p[pre + ".conv2.weight"], None, 1, 1)
        print(pre, "conv2 y", rel_l2(nchw(ws.y[pre + ".conv2"], B, c1.Ho, c1.cout), y))
        y = port._bn(y, p, ob, pre + ".bn2", True)
        if (pre + ".downsample.0.weight") in p:
            d = F.conv2d(h, p[pre + ".downsample.0.weight"], None, s, 0)
            print(pre, "down y", rel_l2(nchw(ws.y[pre + ".downsample.0"], B, c1.Ho, c1.cout), d))
            identity = port._bn(d, p, ob, pre + ".downsample.1", True)
        h = F.relu(y + identity)
        print(pre, "out", rel_l2(nchw(ws.out_f32[pre], B, c1.Ho, c1.cout), h))
print("features", rel_l2(out["features"], torch.flatten(F.adaptive_avg_pool2d(h, (1, 1)), 1)))

# ---- backward: per-tensor gradient errors of one LwF task-0 step -------------------------------------------------------------------------
import libcontinual_b200.model as M
from tests.test_gpu_resnet18 import grads_of
p, b, fc_w, fc_b = synth_resnet18_state(1818, 20)
Bq = int(os.environ.get("DBG_B", "8"))
bb = make_backbone(p, b, max_batch=Bq)
m = M.LWF(bb, 512, 200, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10)
m.before_task(0, None, None, None)
w, bias = m.engine.fc_views(10)
w.copy_(fc_w[:10].cuda()); bias.copy_(fc_b[:10].cuda())
m.train()
x, y = synth_batch(1900, Bq, 0, 10, img=64)
pred, acc, loss = m.observe({"image": x, "label": y})
torch.cuda.synchronize()
got = grads_of(m)
for mode in ("fp32", "bf16"):
    orc = port.ResNetMethodOracle("lwf", p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10, arch="resnet18", maxpool=True, conv_mode=mode)
    _, _, lo, go = orc.step(x, y, apply_update=False)
    print(mode, "loss", float(loss), float(lo))
    for n in go:
        print(f"   {mode} {n:40s} rel {rel_l2(got[n], go[n]):.3f}  norm ours {float(got[n].norm()):.4e} ref {float(go[n].norm()):.4e}")
