"""GPU: DualPrompt on ViT-B/16 (key match, prefix-tuned attention on blocks 0-4, masked CE, backward to the prompt pools) through the C ABI against
tests/golden/dualprompt_vit.npz, written by the REAL reference modules (pool `forward`, ViT blocks with prompt=(pk, pv), classifier; see
oracle/make_golden.py::golden_dualprompt).  Tolerances: BF16 GEMM operands through 12 blocks forward and backward -> 3e-2 relative L2 on the
gradients, 2e-2 on features / loss; fp32 key-match kernel 1e-5; integer selections exact."""
import ctypes

import numpy as np
import pytest
import torch

from tests.golden_util import load, synth_dual_pool, synth_images, synth_vit_state
from tests.test_gpu_kernels import P, dev, lib, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_prompt_key_match_kernel(lib):
    g = torch.Generator().manual_seed(11)
    B, pool, D = 37, 10, 768
    q = torch.randn(B, D, generator=g)
    Ks = [torch.rand(pool, D, generator=g).requires_grad_(True) for _ in range(3)]
    Kd = [dev(k.detach()) for k in Ks]
    dK = [torch.zeros(pool, D, device="cuda") for _ in range(3)]
    Arr = ctypes.c_void_p * 3
    idx = torch.full((3, B), -1, dtype=torch.int64, device="cuda")
    loss = torch.zeros(1, device="cuda")
    assert lib.lc_prompt_key_match(P(dev(q)), Arr(*[P(k) for k in Kd]), Arr(*[P(k) for k in dK]), 3, B, pool, D, 4, P(idx), P(loss), st()) == 0
    torch.cuda.synchronize()
    ref = sum((1.0 - torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(k, dim=1).T)[:, 4].sum() for k in Ks)
    ref.backward()
    assert abs(float(loss) - float(ref)) < 1e-4 * abs(float(ref))
    for k, d in zip(Ks, dK):
        assert rel_l2(d, k.grad) < 1e-5
    assert bool((idx == 4).all())
    # inference: per-sample argmax
    assert lib.lc_prompt_key_match(P(dev(q)), Arr(*[P(k) for k in Kd]), None, 3, B, pool, D, -1, P(idx), None, st()) == 0
    torch.cuda.synchronize()
    for l, k in enumerate(Ks):
        want = (torch.nn.functional.normalize(q, dim=1) @ torch.nn.functional.normalize(k.detach(), dim=1).T).argmax(1)
        assert torch.equal(idx[l].cpu(), want)


def test_gather_rows_bf16(lib):
    g = torch.Generator().manual_seed(2)
    src = torch.rand(10, 20, 768, generator=g)
    idx = torch.randint(0, 10, (9,), generator=g)
    out = torch.zeros(9, 10, 768, dtype=torch.bfloat16, device="cuda")
    sd = dev(src)
    assert lib.lc_gather_rows_bf16(sd.data_ptr() + 4 * 10 * 768, P(dev(idx)), 20 * 768, 10, 768, 9, P(out), st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), src[idx][:, 10:].bfloat16())
    assert lib.lc_gather_rows_bf16(sd.data_ptr(), None, 0, 3, 768, 9, P(out), st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(out.cpu().reshape(-1)[:9 * 3 * 768].reshape(9, 3, 768), src[0, :3].bfloat16().expand(9, -1, -1))


def _model(p, pool, fc_w, fc_b):
    from libcontinual_b200.model import DualPrompt, vit_pt_imnet
    bb = vit_pt_imnet(pretrained=False, state=p, device="cuda:0")
    m = DualPrompt(bb, 768, 100, device="cuda:0", task_num=10, init_cls_num=10, inc_cls_num=10, g_prompt_length=6, e_prompt_length=20)
    with torch.no_grad():
        for k, v in pool.items():
            getattr(m.pool, k).copy_(v.cuda())
    return m


def test_dualprompt_observe_and_inference_match_reference_golden():
    from libcontinual_b200 import optim
    g = load("dualprompt_vit.npz")
    p = synth_vit_state(5150)[0]
    pool, fc_w, fc_b = synth_dual_pool(930)
    m = _model(p, pool, fc_w, fc_b)
    for task in (0, 1):
        m.before_task(task, None, None, None)
        n = m.network.classifier.out_features
        assert n == 10 * (task + 1)
        with torch.no_grad():
            m.head_W[:n].copy_(fc_w[:n].cuda()); m.head_b[:n].copy_(fc_b[:n].cuda())
        lo = 10 * task
        x, y = synth_images(750 + task, 4, lo, lo + 10)
        pred, acc, loss = m.observe({"image": x, "label": y})
        for q in m.get_parameters(None):
            q.grad = None                                # observe -> zero_grad -> backward (trainer.py:601-604)
        loss.backward()
        torch.cuda.synchronize()
        assert not m.engine.tensor_core_error()
        assert abs(float(m.prompt_loss) - float(g[f"t{task}/ploss"])) < 2e-3 * abs(float(g[f"t{task}/ploss"]))
        assert abs(float(loss.detach()) - float(g[f"t{task}/loss"])) < 2e-2 * abs(float(g[f"t{task}/loss"]))
        got = {"dW": m.network.classifier.weight.grad, "db": m.network.classifier.bias.grad}
        got.update({"d" + k: getattr(m.pool, k).grad for k in pool})
        for k, v in got.items():
            e = rel_l2(v, torch.from_numpy(g[f"t{task}/{k}"]))
            print(f"task{task} {k}: rel-L2 vs reference = {e:.2e}")
            assert e < 3e-2, (task, k, e)
        # e-prompt rows of the other tasks and old-class head rows receive exact zeros
        ge = m.pool.e_p_3.grad
        assert float(ge[[i for i in range(10) if i != task]].abs().max()) == 0.0
        assert np.array_equal(pred.cpu().numpy(), g[f"t{task}/pred"])
        # inference: per-sample top-1 key, logits over all seen classes
        ipred, iacc = m.inference({"image": x, "label": y})
        torch.cuda.synchronize()
        bufs = m._batch_bufs(4)
        assert np.array_equal(bufs["idx"].cpu().numpy(), g[f"t{task}/inf_ids"])
        assert rel_l2(bufs["logits"][:, :n], torch.from_numpy(g[f"t{task}/inf_logits"])) < 2e-2
        m.after_task(task, None, None, None)
    # one flat-Adam step == torch.optim.Adam on the same gradients
    m.before_task(2, None, None, None)
    x, y = synth_images(752, 4, 20, 30)
    params = m.get_parameters(None)
    opt = optim.Adam(params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0, model=m)
    ref_params = [q.detach().clone().requires_grad_(True) for q in params]
    ropt = torch.optim.Adam(ref_params, lr=1e-3, betas=(0.9, 0.999), weight_decay=0)
    pred, acc, loss = m.observe({"image": x, "label": y})
    opt.zero_grad()
    loss.backward()
    for rp, q in zip(ref_params, params):
        rp.grad = q.grad.detach().clone()
    opt.step(); ropt.step()
    torch.cuda.synchronize()
    for rp, q in zip(ref_params, params):
        assert rel_l2(q.detach(), rp.detach()) < 1e-6


def test_dualprompt_graphed_step_equals_eager():
    """CUDA-graph replay (GraphedL2PStep without a clip stage) == the eager plugin order with the flat Adam."""
    from libcontinual_b200 import optim
    from libcontinual_b200.trainer import GraphedL2PStep
    p = synth_vit_state(5150)[0]
    pool, fc_w, fc_b = synth_dual_pool(930)
    m = _model(p, pool, fc_w, fc_b)
    m.before_task(0, None, None, None)
    x, y = synth_images(750, 4, 0, 10)
    theta0 = m.theta.clone()
    opt = optim.Adam(m.get_parameters(None), lr=1e-3, model=m)
    for _ in range(2):
        pred, acc, loss = m.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    eager, eager_loss = m.theta.clone(), float(loss.detach())
    m.theta.copy_(theta0)
    opt2 = optim.Adam(m.get_parameters(None), lr=1e-3, model=m)
    gs = GraphedL2PStep(m, opt2, 4)
    for _ in range(2):
        gs.run(x, y)
    torch.cuda.synchronize()
    assert rel_l2(m.theta, eager) < 1e-6 and abs(float(gs.loss()) - eager_loss) < 1e-5
