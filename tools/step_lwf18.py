"""One eager LwF / ResNet18 step at batch 256 (task 1, teacher live) — the target of an `ncu` launch list (tools/launch_summary.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
m, lo, hi = bench.build_model("lwf18", torch.device("cuda", 0))
from libcontinual_b200.optim import SGD
opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=m.engine)
x, y = bench.synth_batches(1, hi, lo, batch=256, img=64)[0]
x, y = x.cuda(), y.cuda()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    m._launch_step(x, y)
    m.engine.sgd_step(opt.buf, opt.hp)
torch.cuda.synchronize()
print("ok", float(m.engine.scal[0]))
