"""Where a ResNet32 step's time goes: CUDA-graph replays of (a) the backbone forward alone, (b) forward + head + loss, (c) the backbone backward
alone, (d) the whole step, at batch 128 in tensor-core mode.  Run once per variant:
    LC_RESNET_UNFUSED=1  (stand-alone BatchNorm-backward launches)   LC_RESNET_SERIAL=1  (weight gradients on the main chain)
Usage: python tools/step_breakdown.py [ewc|icarl] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench

wl = sys.argv[1] if len(sys.argv) > 1 else "ewc"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
dev = torch.device("cuda", 0)
m, lo, hi = bench.build_model(wl, dev)
eng = m.engine
from libcontinual_b200.optim import SGD

opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=eng)
x, y = bench.synth_batches(1, hi, lo)[0]
x, y = x.cuda(), y.cuda()


def timed(fn, name):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(10):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:44s} {e0.elapsed_time(e1) / reps * 1e3:9.1f} us", flush=True)


def full():
    m._launch_step(x, y)
    eng.sgd_step(opt.buf, opt.hp)


print(f"workload {wl}  UNFUSED={os.environ.get('LC_RESNET_UNFUSED', '0')} SERIAL={os.environ.get('LC_RESNET_SERIAL', '0')}")
timed(lambda: eng.forward(x, train=True, update_running=False), "backbone forward (train)")
timed(lambda: eng.forward(x, train=False, update_running=False), "backbone forward (eval: teacher)")
full(); torch.cuda.synchronize()
timed(lambda: eng.backward(x), "backbone backward")
timed(full, "whole step (fwd + loss + bwd + sgd)")
