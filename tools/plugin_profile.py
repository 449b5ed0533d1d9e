"""Host-side cost of the reference step order on the plugin surface (observe -> zero_grad -> backward -> step -> loss.item()), per phase."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from libcontinual_b200.optim import SGD
wl = sys.argv[1] if len(sys.argv) > 1 else "icarl"
m, lo, hi = bench.build_model(wl, torch.device("cuda", 0))
opt = SGD(m.get_parameters(None), lr=0.1, momentum=0.9, weight_decay=5e-4, engine=m.engine)
host = [(x.pin_memory(), y.pin_memory()) for x, y in bench.synth_batches(4, hi, lo)]
T = {k: 0.0 for k in ("observe", "zero_grad", "backward", "step", "item")}
N = 100
for i in range(N + 10):
    if i == 10:
        T = {k: 0.0 for k in T}
        torch.cuda.synchronize(); t_all = time.perf_counter()
    x, y = host[i % 4]
    t0 = time.perf_counter(); pred, acc, loss = m.observe({"image": x, "label": y})
    t1 = time.perf_counter(); opt.zero_grad()
    t2 = time.perf_counter(); loss.backward()
    t3 = time.perf_counter(); opt.step()
    t4 = time.perf_counter(); v = loss.item()
    t5 = time.perf_counter()
    for k, d in zip(T, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4)):
        T[k] += d
torch.cuda.synchronize()
tot = (time.perf_counter() - t_all) / N * 1e3
print(f"{wl}: {tot:.3f} ms / step  ->  {128 / tot * 1e3:.0f} img/s")
for k, v in T.items():
    print(f"  {k:10s} {v / N * 1e3:7.3f} ms")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for i in range(50):
    x, y = host[i % 4]
    pred, acc, loss = m.observe({"image": x, "label": y}); opt.zero_grad(); loss.backward(); opt.step(); loss.item()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
