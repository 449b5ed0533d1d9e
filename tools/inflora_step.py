"""Runs a few eager InfLoRA_OPT steps at bs=128 (for `ncu` launch lists / captures): python tools/inflora_step.py [steps] [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from libcontinual_b200.model.inflora import InfLoRA_OPT
from libcontinual_b200.model.l2p import vit_pt_imnet
from libcontinual_b200.optim import FlatSGD

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 128
os.environ.setdefault("PYTHONHASHSEED", "42")
p, A, Bm, hw, hb = bench.inflora_synth_state()
bb = vit_pt_imnet(pretrained=False, state=p, device="cuda:0", attn_layer="MultiHeadAttention_LoRA", lora_rank=10)
m = InfLoRA_OPT(bb, "cuda:0", init_cls_num=20, inc_cls_num=20, task_num=10, lame=1.0, lamb=0.95, embd_dim=768, use_ca=False, dataset="imagenet-r")
m.start_task(0); m.start_task(1, A)
with torch.no_grad():
    m.lora_B.copy_(Bm.cuda()); m.heads_W.copy_(hw.cuda()); m.heads_b.copy_(hb.cuda())
opt = FlatSGD(m.get_parameters(None), lr=8e-3, momentum=0.9, model=m)
x, y = bench.l2p_batches(1, batch, 20, 40)[0]
x, y = x.cuda(), y.cuda()
torch.cuda.synchronize()
for i in range(steps):
    pred, acc, loss = m.observe({"image": x, "label": y})
    opt.zero_grad()
    loss.backward()
    opt.step()
    print("step", i, "loss", float(loss), "acc", acc, flush=True)
print("tc error:", m.engine.tensor_core_error())
