"""Runs a few eager L2P steps at bs=128 (for `ncu` launch lists / captures): python tools/l2p_step.py [steps] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from libcontinual_b200.model.l2p import L2P, vit_pt_imnet
from libcontinual_b200.optim import Adam

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
p, prm, key, fc_w, fc_b = bench.l2p_synth_state()
bb = vit_pt_imnet(pretrained=False, state=p, device="cuda:0")
m = L2P(bb, "cuda:0", init_cls_num=10, inc_cls_num=10, num_class=100, task_num=10, feat_dim=768, prompt_length=5, pool_size=10, top_k=5, pull_constraint_coeff=1.0)
with torch.no_grad():
    bb.prompt.prompt.copy_(prm); bb.prompt.prompt_key.copy_(key)
opt = Adam(m.get_parameters(None), lr=0.001875, model=m)
x, y = bench.l2p_batches(1, batch, 0, 10)[0]
x, y = x.cuda(), y.cuda()
torch.cuda.synchronize()
for i in range(steps):
    opt.zero_grad()
    pred, acc, loss = m.observe({"image": x, "label": y})
    opt.step()
    print("step", i, "loss", float(loss), "acc", acc, flush=True)
print("tc error:", m.engine.tensor_core_error())
