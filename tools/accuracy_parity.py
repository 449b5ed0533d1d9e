"""Final average accuracy of a short synthetic class-incremental stream (EWC, cifar_resnet32, 2 tasks x 10 classes, bs 32, SGD 0.1/0.9/5e-4, lamda 1000:
config/ewc.yaml with BASELINE config C1's overrides) on the CUDA path (precision 'tc' and 'fp32') next to the CPU oracle, from identical initial
weights and identical batches.  The data have learnable structure: image = amp * template[class] + N(0, 1) noise.
With lr 0.1 / lamda 1000 such short streams are chaotic, so the tool runs MANY seeds and reports, per arm, mean +- std of the final average accuracy,
the paired difference to the oracle (mean +- standard error), and a CONTROL arm — the oracle itself restarted from weights moved by one fp32 ulp — whose
paired difference to the oracle is the band inside which no two faithful implementations can be told apart.

    python tools/accuracy_parity.py [steps_per_task] [n_test_per_task] [--seeds N]   -> one JSON line
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def make_stream(seed, steps, n_test, bs=32, n_tasks=2, cpt=10, amp=1.0):
    rng = np.random.default_rng(seed)
    templ = rng.standard_normal((n_tasks * cpt, 3, 32, 32)).astype(np.float32)

    def draw(n, lo, hi):
        y = rng.integers(lo, hi, (n,))
        x = amp * templ[y] + rng.standard_normal((n, 3, 32, 32)).astype(np.float32)
        return torch.from_numpy(x), torch.from_numpy(y.astype(np.int64))
    train = [[draw(bs, t * cpt, (t + 1) * cpt) for _ in range(steps)] for t in range(n_tasks)]
    test = [draw(n_test, t * cpt, (t + 1) * cpt) for t in range(n_tasks)]
    return train, test


METHOD, LR, AMP = "ewc", 0.1, 1.0


def run_ours(p, b, fc_w, fc_b, train, test, precision):
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    bb = M.cifar_resnet32(max_batch=256, precision=precision)
    bb.load_state_dict({**p, **b}, strict=True)
    if METHOD == "lwf":
        return run_ours_lwf(bb, fc_w, fc_b, train, test)
    m = M.EWC(bb, 64, 100, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10, lamda=1000.0)

    class Loader(list):
        batch_size = 32
    for t, batches in enumerate(train):
        m.before_task(t, None, None, None)
        n = 10 * (t + 1)
        w, bias = m.engine.fc_views(n)
        w[10 * t:].copy_(fc_w[10 * t:n].cuda()); bias[10 * t:].copy_(fc_b[10 * t:n].cuda())
        m.train()
        opt = SGD(m.get_parameters(None), lr=LR, momentum=0.9, weight_decay=5e-4, engine=m.engine)
        for x, y in batches:
            pred, acc, loss = m.observe({"image": x, "label": y})
            opt.zero_grad(); loss.backward(); opt.step()
        m.after_task(t, None, Loader([{"image": x, "label": y} for x, y in batches]), None)
    return eval_ours(m, test)


def eval_ours(m, test):
    m.eval()
    accs = []
    for x, y in test:
        ok = 0
        for i in range(0, x.shape[0], 250):
            _, a = m.inference({"image": x[i:i + 250], "label": y[i:i + 250]})
            ok += a * min(250, x.shape[0] - i)
        accs.append(ok / x.shape[0])
    return accs


def run_ours_lwf(bb, fc_w, fc_b, train, test):
    """LwF (lwf.py:52-70: CE on the new slice + 3 * KD(T=2) against the frozen previous model), same stream."""
    import libcontinual_b200.model as M
    from libcontinual_b200.optim import SGD
    m = M.LWF(bb, 64, 100, device=torch.device("cuda"), init_cls_num=10, inc_cls_num=10)
    for t, batches in enumerate(train):
        m.before_task(t, None, None, None)
        n = 10 * (t + 1)
        w, bias = m.engine.fc_views(n)
        w[10 * t:].copy_(fc_w[10 * t:n].cuda()); bias[10 * t:].copy_(fc_b[10 * t:n].cuda())
        m.train()
        opt = SGD(m.get_parameters(None), lr=LR, momentum=0.9, weight_decay=5e-4, engine=m.engine)
        for x, y in batches:
            pred, acc, loss = m.observe({"image": x, "label": y})
            opt.zero_grad(); loss.backward(); opt.step()
    return eval_ours(m, test)


def run_oracle(p, b, fc_w, fc_b, train, test):
    from oracle import port
    torch.set_num_threads(os.cpu_count() or 1)
    orc = port.ResNetMethodOracle(METHOD, p, b, fc_w[:10], fc_b[:10], init_cls=10, inc_cls=10, lamda=1000.0, lr=LR)
    for t, batches in enumerate(train):
        if t > 0:
            if METHOD == "lwf":
                orc.snapshot_teacher(); orc.prev_cls = 10 * t
            orc.task_idx = t
            orc.grow_head(fc_w[:10 * (t + 1)], fc_b[:10 * (t + 1)]); orc.reset_optimizer()
        for x, y in batches:
            orc.step(x, y)
        if METHOD == "ewc":
            orc.ewc_after_task(batches, 32)
    accs = []
    with torch.no_grad():
        for x, y in test:
            pred = torch.cat([orc.logits(x[i:i + 500], False).argmax(1) for i in range(0, x.shape[0], 500)])
            accs.append(float((pred == y).float().mean()))
    return accs


def one_seed(seed, steps, n_test):
    from oracle import port
    rng = np.random.default_rng(1000 + seed)
    p, b = port.cifar_resnet_init(rng)
    bound = 1.0 / 8.0
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (20, 64)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (20,)).astype(np.float32))
    train, test = make_stream(2000 + seed, steps, n_test, amp=AMP)
    # control arm: the SAME oracle from initial weights moved by one fp32 ulp (relative 6e-8) — how far the reference's final accuracy moves under a
    # perturbation no implementation can be asked to reproduce
    prng = np.random.default_rng(3000 + seed)
    p_eps = {k: v * torch.from_numpy(1 + 6e-8 * prng.standard_normal(tuple(v.shape))).float() for k, v in p.items()}
    out = {}
    for name, fn in (("cuda_tc", lambda: run_ours(p, b, fc_w, fc_b, train, test, "tc")), ("cuda_fp32", lambda: run_ours(p, b, fc_w, fc_b, train, test, "fp32")),
                     ("oracle_cpu_fp32", lambda: run_oracle(p, b, fc_w, fc_b, train, test)),
                     ("oracle_cpu_fp32_1ulp", lambda: run_oracle(p_eps, b, fc_w, fc_b, train, test))):
        out[name] = 100 * float(np.mean(fn()))
    return out


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("steps", nargs="?", type=int, default=60)
    ap.add_argument("n_test", nargs="?", type=int, default=1000)
    ap.add_argument("--seeds", type=int, default=1)
    ap.add_argument("--method", default="ewc", choices=["ewc", "lwf"])
    ap.add_argument("--lr", type=float, default=0.1)
    ap.add_argument("--amp", type=float, default=1.0, help="class-template amplitude against unit noise")
    ap.add_argument("--oracle-only", action="store_true", help="CPU: only the oracle and its 1-ulp control arm (stream design)")
    ap.add_argument("--ours-only", action="store_true", help="GPU box: only the two CUDA arms (the oracle arms of the same seeds run on a CPU-only machine)")
    ap.add_argument("--merge", nargs=2, metavar=("OURS_JSON", "ORACLE_JSON"), help="combine an --ours-only and an --oracle-only run of the same seeds")
    a = ap.parse_args()
    global METHOD, LR, AMP
    METHOD, LR, AMP = a.method, a.lr, a.amp
    if a.oracle_only:
        global run_ours
        run_ours = lambda *args, **kw: [float("nan")]
    if a.ours_only:
        global run_oracle
        run_oracle = lambda *args, **kw: [float("nan")]
    if a.merge:
        ours, orc = (json.load(open(f)) for f in a.merge)
        assert len(ours["per_seed"]) == len(orc["per_seed"])
        rows = [{k: (o[k] if k.startswith("cuda") else r[k]) for k in o} for o, r in zip(ours["per_seed"], orc["per_seed"])]
        a.seeds = len(rows)
    else:
        rows = [one_seed(s, a.steps, a.n_test) for s in range(a.seeds)]
    arms = list(rows[0])
    res = {"stream": f"{METHOD.upper()} cifar_resnet32, 2 tasks x 10 classes, {a.steps} steps/task, bs 32, SGD {LR}/0.9/5e-4, "
                     + ("lamda 1000, " if METHOD == "ewc" else "KD weight 3 T 2, ") + f"template amplitude {AMP}, test {a.n_test}/task "
                     f"(synthetic class templates + N(0,1) noise); {a.seeds} seeds (weights and data re-drawn per seed); metric = final average accuracy (%)",
           "per_seed": rows}
    for arm in arms:
        v = np.array([r[arm] for r in rows])
        res[arm] = {"mean": round(float(v.mean()), 3), "std": round(float(v.std(ddof=1)) if len(v) > 1 else 0.0, 3)}
    base = np.array([r["oracle_cpu_fp32"] for r in rows])
    for arm in arms:
        if arm == "oracle_cpu_fp32":
            continue
        d = np.array([r[arm] for r in rows]) - base
        res["delta_pp_" + arm + "_vs_oracle"] = {"mean": round(float(d.mean()), 3), "stderr": round(float(d.std(ddof=1) / np.sqrt(len(d))) if len(d) > 1 else 0.0, 3),
                                                  "mean_abs": round(float(np.abs(d).mean()), 3)}
    print(json.dumps(res))
    return res


if __name__ == "__main__":
    main()
