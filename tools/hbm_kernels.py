"""Launch every coalesced / vectorised HBM kernel north_star names at its benchmark size, a few times each, for an ncu capture:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum --clock-control none \
        -k regex:'ewc_penalty|fisher_|sgd_momentum|adam_kernel|ce_kd_loss|cosine_head|lucir_loss|l2p_select|gpm_project|sqnorm|clip_scale' \
        --csv --log-file gpurun_out/hbm.csv python tools/hbm_kernels.py
    python tools/hbm_summary.py gpurun_out/hbm.csv

EWC penalty + gradient, Fisher accumulate / merge, SGD momentum over the 472 k-float ResNet32 arena (bs 128); the CE + KD loss kernel (iCaRL: 55 classes,
50 distilled); LUCIR's cosine head forward / backward + loss; the CUDA-core GPM projection at the AlexNet_TRGP layer shapes.  Adam + clip + l2p_select,
the LoRA merge and the tcgen05 GPM projection are captured from whole steps of `bench.py --workload l2p | inflora | gpm`."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
import libcontinual_b200.model as M
from libcontinual_b200.optim import SGD

dev = torch.device("cuda", 0)
REPS = 3

# ---- EWC: penalty + gradient, Fisher accumulate / merge, SGD, CE loss ------------------------------------------------------------------------
m, lo, hi = bench.build_model("ewc", dev)
eng = m.engine
opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=eng)
x, y = bench.synth_batches(1, hi, lo)[0]
x, y = x.cuda(), y.cuda()
for _ in range(REPS):
    m._launch_step(x, y)                      # forward, CE on the new slice, backward, ewc_penalty_grad_kernel
    eng.sgd_step(opt.buf, opt.hp)             # sgd_momentum_kernel
fisher = torch.zeros_like(eng.params)
for _ in range(REPS):
    eng.fisher_accumulate(fisher, 128.0)      # fisher_accumulate_kernel
from libcontinual_b200._lib import check, ptr, stream_ptr

for _ in range(REPS):
    check(eng.lib.lc_fisher_merge(ptr(fisher), ptr(m.fisher), eng.n_total, 5000.0, 0.5, stream_ptr()))
torch.cuda.synchronize()
print("ewc ok", float(eng.scal[0]))
del m, eng, opt

# ---- iCaRL: CE + KD loss kernel ------------------------------------------------------------------------------------------------------------------
m, lo, hi = bench.build_model("icarl", dev)
x, y = bench.synth_batches(1, hi, lo)[0]
x, y = x.cuda(), y.cuda()
for _ in range(REPS):
    m._launch_step(x, y)                      # ce_kd_loss_kernel with the distillation term live
torch.cuda.synchronize()
print("icarl ok", float(m.engine.scal[0]))
del m

# ---- LUCIR: cosine head forward / backward + less-forget / margin-ranking loss ---------------------------------------------------------------------
try:
    bb = M.resnet32_V2(max_batch=128)
    m = M.LUCIR(bb, 64, 100, device=dev, init_cls_num=50, inc_cls_num=10, K=2, lw_mr=1, lamda=5, dist=0.5)
    m.before_task(0, None, None, None)
    m.train()
    xs, ys = bench.synth_batches(2, 60, 0)[0]
    pred, acc, loss = m.observe({"image": xs, "label": ys.clamp(max=49)})
    m.before_task(1, None, None, None)        # 60 classes, frozen reference model = the task-0 network
    for _ in range(REPS):
        pred, acc, loss = m.observe({"image": xs, "label": ys})
        loss.backward()
    torch.cuda.synchronize()
    print("lucir ok", float(loss))
    del m, bb
except Exception as e:                        # keep the remaining captures
    print("lucir section failed:", repr(e))

# ---- GPM projection (CUDA-core fp32 kernel) at the AlexNet_TRGP layer shapes (gpm.py:78-81) --------------------------------------------------------------
lib = __import__("libcontinual_b200._lib", fromlist=["load"]).load()
for cout, d in [(64, 48), (128, 576), (256, 512), (2048, 1024), (2048, 2048)]:
    g = torch.randn(cout, d, device=dev)
    Mm = torch.randn(d, d, device=dev) * 0.01
    for _ in range(REPS):
        check(lib.lc_gpm_project(ptr(g), ptr(Mm), cout, d, stream_ptr()))
torch.cuda.synchronize()
print("gpm ok")
# Adam / clip / l2p_select, lora_merge and the tcgen05 GPM projection are captured from whole steps: `bench.py --workload l2p|inflora|gpm --only --steps 2`.
