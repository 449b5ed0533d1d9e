import sys; sys.path.insert(0, '.')
import numpy as np, torch
from oracle import port
from tests.golden_util import synth_batch, synth_resnet_state
from tests.test_gpu_resnet import make_backbone, load_head, grads_of, rel_l2
import libcontinual_b200.model as M
B = 8
p, b, fc_w, fc_b = synth_resnet_state(101, 20)
bb = make_backbone(p, b)
m = M.EWC(bb, 64, 100, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10, lamda=1000.0)
m.before_task(0, None, None, None); load_head(m, fc_w[:10], fc_b[:10]); m.train()
orc = port.ResNetMethodOracle("ewc", p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10, lamda=1000.0)
x, y = synth_batch(1000, B, 0, 10)
po, ao, lo, go = orc.step(x, y, apply_update=False)
for rep in range(3):
    m.engine.rstat.copy_(torch.zeros_like(m.engine.rstat)); m.engine.reset_running_stats()
    pred, acc, loss = m.observe({"image": x, "label": y})
    torch.cuda.synchronize()
    got = grads_of(m)
    errs = {n: rel_l2(got[n], go[n]) for n in go}
    bad = sorted(errs.items(), key=lambda kv: -kv[1])[:6]
    print(rep, float(loss), float(lo), bad)
    n = bad[0][0]
    print('   got', got[n].flatten()[:8].tolist()); print('   ref', go[n].flatten()[:8].tolist())
    nb = n.replace('.weight', '.bias')
    if nb in got: print('   bias got', got[nb].flatten()[:6].tolist(), 'ref', go[nb].flatten()[:6].tolist())
print("---- backward path")
pred, acc, loss = m.observe({"image": x, "label": y})
torch.cuda.synchronize()
before = m.engine.grads.clone()
loss.backward()
torch.cuda.synchronize()
after = m.engine.grads
print("arena changed by backward:", float((after - before).abs().max()))
eng = m.engine
for (n, _), q in zip(eng.layout, [q for _, q in m.backbone.named_parameters()]):
    a = eng.param_view(n, before)
    d = float((q.grad - a).abs().max())
    inarena = eng.grads.data_ptr() <= q.grad.data_ptr() < eng.grads.data_ptr() + 4 * eng.grads.numel()
    if d > 0 or not inarena:
        print("  ", n, "p.grad vs arena diff", d, "grad in arena:", inarena, q.grad.data_ptr() - eng.grads.data_ptr(), eng.param_off[n][0] * 4)
got = grads_of(m)
errs = {n: rel_l2(got[n], go[n]) for n in go}
print("worst after backward", sorted(errs.items(), key=lambda kv: -kv[1])[:3])
# second step without zeroing p.grad
pred, acc, loss = m.observe({"image": x, "label": y})
loss.backward(); torch.cuda.synchronize()
got = grads_of(m)
errs = {n: rel_l2(got[n], go[n]) for n in go}
print("worst after 2nd backward (no zero_grad)", sorted(errs.items(), key=lambda kv: -kv[1])[:3])
