"""GPU: CodaPrompt on ViT-B/16 (attention-weighted prompt kernel, prefix-tuned attention on blocks 0-4, backward to components / keys / attention
vectors) through the C ABI against tests/golden/codaprompt_vit.npz, written by the REAL reference modules (oracle/make_golden.py::golden_codaprompt).
Tolerances: fp32 pool kernels 1e-5 vs float64; network level (BF16 GEMM operands through 12 blocks) 3e-2 relative L2 on gradients, 2e-2 on logits."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import port
from tests.golden_util import load, synth_coda_pool, synth_images, synth_vit_state
from tests.test_gpu_kernels import P, dev, lib, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_coda_prompt_kernels(lib):
    g = torch.Generator().manual_seed(21)
    B, pool_n, nk, Lp, D, nl = 19, 12, 10, 8, 768, 2
    q = torch.randn(B, D, generator=g)
    K = [torch.randn(pool_n, D, generator=g) for _ in range(nl)]
    A = [torch.randn(pool_n, D, generator=g) for _ in range(nl)]
    pp = [torch.randn(pool_n, Lp, D, generator=g) * 0.1 for _ in range(nl)]
    Arr = ctypes.c_void_p * nl
    Kd, Ad, pd = [dev(t) for t in K], [dev(t) for t in A], [dev(t) for t in pp]
    pk = [torch.zeros(B, Lp // 2, D, dtype=torch.bfloat16, device="cuda") for _ in range(nl)]
    pv = [torch.zeros(B, Lp // 2, D, dtype=torch.bfloat16, device="cuda") for _ in range(nl)]
    alpha = torch.zeros(nl, B, nk, device="cuda"); vnorm = torch.zeros(nl, B, nk, device="cuda"); dalpha = torch.zeros(nl, B, nk, device="cuda")
    qd = dev(q)
    ptrs = lambda ts: Arr(*[P(t) for t in ts])
    assert lib.lc_coda_prompt_forward(P(qd), ptrs(Kd), ptrs(Ad), ptrs(pd), ptrs(pk), ptrs(pv), nl, B, nk, Lp, D, P(alpha), P(vnorm), st()) == 0
    torch.cuda.synchronize()
    Kr = [t.double().requires_grad_(True) for t in K]; Ar = [t.double().requires_grad_(True) for t in A]; pr = [t.double().requires_grad_(True) for t in pp]
    pool = {}
    for l in range(nl):
        pool[f"e_k_{l}"], pool[f"e_a_{l}"], pool[f"e_p_{l}"] = Kr[l], Ar[l], pr[l]
    for l in range(nl, 5):
        pool[f"e_k_{l}"], pool[f"e_a_{l}"], pool[f"e_p_{l}"] = Kr[0], Ar[0], pr[0]
    ref = port.codaprompt_prefixes(pool, q.double(), nk)
    for l in range(nl):
        assert rel_l2(pk[l].float(), ref[l][0]) < 4e-3 and rel_l2(pv[l].float(), ref[l][1]) < 4e-3      # BF16 output rounding
    # backward from random prefix-row gradients
    dpk = [torch.randn(B, Lp // 2, D, generator=g) for _ in range(nl)]; dpv = [torch.randn(B, Lp // 2, D, generator=g) for _ in range(nl)]
    dK = [torch.zeros(pool_n, D, device="cuda") for _ in range(nl)]; dA = [torch.zeros(pool_n, D, device="cuda") for _ in range(nl)]
    dp = [torch.zeros(pool_n, Lp, D, device="cuda") for _ in range(nl)]
    assert lib.lc_coda_prompt_backward(P(qd), ptrs(Kd), ptrs(Ad), ptrs(pd), ptrs([dev(t) for t in dpk]), ptrs([dev(t) for t in dpv]), ptrs(dK), ptrs(dA), ptrs(dp),
                                       nl, B, nk, Lp, D, P(alpha), P(vnorm), P(dalpha), st()) == 0
    torch.cuda.synchronize()
    tot = sum((ref[l][0] * dpk[l].double()).sum() + (ref[l][1] * dpv[l].double()).sum() for l in range(nl))
    tot.backward()
    for l in range(nl):
        assert rel_l2(dK[l][:nk], Kr[l].grad[:nk]) < 1e-5, l
        assert rel_l2(dA[l][:nk], Ar[l].grad[:nk]) < 1e-5, l
        assert rel_l2(dp[l][:nk], pr[l].grad[:nk]) < 1e-5, l
        assert float(dK[l][nk:].abs().max()) == 0.0


def test_codaprompt_observe_and_inference_match_reference_golden():
    from libcontinual_b200.model import CodaPrompt, vit_pt_imnet
    g = load("codaprompt_vit.npz")
    p = synth_vit_state(5150)[0]
    pool, fc_w, fc_b = synth_coda_pool(940)
    bb = vit_pt_imnet(pretrained=False, state=p, device="cuda:0")
    m = CodaPrompt(bb, 768, 100, device="cuda:0", task_num=10, init_cls_num=10, inc_cls_num=10, prompt_length=8, pool_size=100, mu=0.0)
    with torch.no_grad():
        for k, v in pool.items():
            getattr(m.pool, k).copy_(v.cuda())
    for task in (0, 1):
        m.before_task(task, None, None, None)
        n = m.network.classifier.out_features
        with torch.no_grad():
            m.head_W[:n].copy_(fc_w[:n].cuda()); m.head_b[:n].copy_(fc_b[:n].cuda())
        lo = 10 * task
        x, y = synth_images(760 + task, 4, lo, lo + 10)
        pred, acc, loss = m.observe({"image": x, "label": y})
        for q in m.get_parameters(None):
            q.grad = None
        loss.backward()
        torch.cuda.synchronize()
        assert not m.engine.tensor_core_error()
        assert abs(float(loss.detach()) - float(g[f"t{task}/loss"])) < 2e-2 * abs(float(g[f"t{task}/loss"]))
        got = {"dW": m.network.classifier.weight.grad, "db": m.network.classifier.bias.grad}
        got.update({"d" + k: getattr(m.pool, k).grad[:10] for k in pool})
        for k, v in got.items():
            e = rel_l2(v, torch.from_numpy(g[f"t{task}/{k}"]))
            print(f"task{task} {k}: rel-L2 vs reference = {e:.2e}")
            assert e < 3e-2, (task, k, e)
        assert float(m.pool.e_p_2.grad[10:].abs().max()) == 0.0            # unused components: exact zeros
        assert np.array_equal(pred.cpu().numpy(), g[f"t{task}/pred"])
        ipred, iacc = m.inference({"image": x, "label": y})
        torch.cuda.synchronize()
        assert rel_l2(m._batch_bufs(4)["logits"][:, :n], torch.from_numpy(g[f"t{task}/inf_logits"])) < 2e-2
        m.after_task(task, None, None, None)
