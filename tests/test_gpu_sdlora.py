"""GPU: SD-LoRA on ViT-B/16 (scaled multi-adapter merge, rank-form dA / dB, magnitude gradients) through the C ABI against tests/golden/sdlora_vit.npz,
written by the REAL `core.model.sd_lora.SD_LoRA.observe` + backward.  Tolerances: fp32 kernels 1e-5 vs float64; network level (BF16 GEMM operands
through 12 blocks) 3e-2 relative L2 on gradients, 2e-2 on logits / loss."""
import numpy as np
import pytest
import torch

from tests.golden_util import load, sdlora_task_states, synth_images, synth_vit_state
from tests.test_gpu_kernels import P, dev, lib, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_rowouter_transposed_scaled_and_coldot(lib):
    g = torch.Generator().manual_seed(5)
    n, D, r, R = 777, 768, 10, 30
    X = (torch.randn(n, D, generator=g) * 0.3).bfloat16()
    Z = torch.randn(n, 64, generator=g)
    nchunk = 23
    partial = torch.zeros(lib.lc_lora_bgrad_partial_floats(2, D, r, nchunk), device="cuda")
    out = torch.full((2, r, D), float("nan"), device="cuda")
    sc = torch.tensor([0.7], device="cuda")
    # two "slabs" reading the same X (x_slab_stride 0), Z columns z0 + s*32 + j, transposed output [s][j][c], device scale
    assert lib.lc_rowouter_bf16(P(dev(X)), D, 0, 0, 2, D, P(dev(Z)), 64, 20, 32, r, n, P(partial), nchunk, P(out), 1, P(sc), st()) == 0
    torch.cuda.synchronize()
    ref = torch.stack([0.7 * (Z[:, 20:30].double().T @ X.double()), 0.7 * (Z[:, 52:62].double().T @ X.double())])
    assert rel_l2(out, ref) < 1e-5
    # magnitude gradient: dmag[i] += sum_jj w[i*r+jj] sum_n G[n][i*r+jj] Z[n][z0+i*r+jj]
    G = torch.randn(n, 32, generator=g)
    w = torch.rand(R, generator=g)
    dmag = torch.tensor([1.0, 2.0, 3.0], device="cuda")
    cd = torch.zeros(50 * R, device="cuda")
    assert lib.lc_coldot_accumulate(P(dev(G)), 32, P(dev(Z)), 64, 4, R, r, n, P(dev(w)), P(cd), 50, P(dmag), st()) == 0
    torch.cuda.synchronize()
    col = (G[:, :R].double() * Z[:, 4:4 + R].double()).sum(0) * w.double()
    ref = torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64) + col.reshape(3, r).sum(1)
    assert rel_l2(dmag, ref) < 1e-5


def test_lora_merge_stacked_scaled(lib):
    """W' = W + B diag(s) A with three adapters stacked along the rank axis (InfLoRA-orig's sum over tasks, vit_inflora.py:236-240, is the s = 1 case)."""
    g = torch.Generator().manual_seed(8)
    L, D, R = 1, 768, 30
    W = torch.randn(L, 3 * D, D, generator=g) * 0.02
    A = torch.randn(L, 2, R, D, generator=g) * 0.03
    B = torch.randn(L, 2, D, R, generator=g) * 0.05
    s = torch.rand(L, 2, R, generator=g)
    wout = torch.zeros(L, 3 * D, D, device="cuda")
    assert lib.lc_lora_merge(P(dev(W)), P(dev(A)), P(dev(B)), P(dev(s)), 0b101, L, D, R, None, None, P(wout), st()) == 0
    torch.cuda.synchronize()
    ref_q = W[0, :D].double() + (B[0, 0].double() * s[0, 0].double()) @ A[0, 0].double()
    ref_v = W[0, 2 * D:].double() + (B[0, 1].double() * s[0, 1].double()) @ A[0, 1].double()
    assert (wout[0, :D].cpu().double() - ref_q).abs().max().item() < 1e-6 and (wout[0, 2 * D:].cpu().double() - ref_v).abs().max().item() < 1e-6
    assert wout[0, D:2 * D].abs().max().item() == 0


def _model(p):
    from libcontinual_b200.model import SD_LoRA, vit_pt_imnet
    bb = vit_pt_imnet(pretrained=False, state=p, device="cuda:0", attn_layer="MultiHeadAttention_SDLoRA", lora_rank=10)
    return SD_LoRA(bb, "cuda:0", init_cls_num=10, inc_cls_num=10, task_num=10, embd_dim=768, init_mag=1.0, rank_reduction=[False, 4, 8, 8, 6],
                   knowledge_dist=[False, 9e-4], dataset="cifar100")


def _install(m, task, blocks, mags, hw, hb):
    with torch.no_grad():
        for l in range(12):
            ad = blocks[l][task]
            m.A_cur[l, 0].copy_(ad["A_q"].cuda()); m.A_cur[l, 1].copy_(ad["A_v"].cuda())
            m.B_cur[l, 0].copy_(ad["B_q"].cuda()); m.B_cur[l, 1].copy_(ad["B_v"].cuda())
        m.mag_all[:task + 1].copy_(mags.cuda())
        m.head_W[:hw.shape[0]].copy_(hw.cuda()); m.head_b[:hb.shape[0]].copy_(hb.cuda())


def test_sdlora_observe_matches_reference_golden():
    from libcontinual_b200 import optim
    g = load("sdlora_vit.npz")
    p = synth_vit_state(5150)[0]
    m = _model(p)
    states = sdlora_task_states(2)
    for task in (0, 1, 2):
        blocks, mags, hw, hb = states[task]
        m.before_task(task, None, None, None)
        _install(m, task, blocks, mags, hw, hb)
        if task == 1:
            m.after_task(task, None, None, None)
            continue
        lo = 10 * task
        x, y = synth_images(780 + task, 4, lo, lo + 10)
        pred, acc, loss = m.observe({"image": x, "label": y})
        for q in m.get_parameters(None):
            q.grad = None
        loss.backward()
        torch.cuda.synchronize()
        assert not m.engine.tensor_core_error()
        assert abs(float(loss.detach()) - float(g[f"t{task}/loss"])) < 2e-2 * abs(float(g[f"t{task}/loss"]))
        head = m._network.classifier
        got = {"dW": head.weight.grad, "db": head.bias.grad, "dmag": torch.cat([q.grad for q in m.mag_lora]),
               "dA_q": torch.stack([q.grad for q in m.lora_A_q]), "dA_v": torch.stack([q.grad for q in m.lora_A_v]),
               "dB_q": torch.stack([q.grad for q in m.lora_B_q]), "dB_v": torch.stack([q.grad for q in m.lora_B_v])}
        for k, v in got.items():
            e = rel_l2(v, torch.from_numpy(g[f"t{task}/{k}"]))
            print(f"task{task} {k}: rel-L2 vs reference = {e:.2e}")
            assert e < 3e-2, (task, k, e)
        n = 10 * (task + 1)
        assert rel_l2(m._batch_bufs(4)["logits"][:, :n], torch.from_numpy(g[f"t{task}/logits"])) < 2e-2
        assert np.array_equal(pred.cpu().numpy(), g[f"t{task}/pred"])
        m.after_task(task, None, None, None)
    # FlatSGD over the active ranges == torch.optim.SGD on the same gradients (task 3, fresh adapter)
    m.before_task(3, None, None, None)
    params = m.get_parameters(None)
    opt = optim.FlatSGD(params, lr=8e-3, momentum=0.9, model=m)
    ref_params = [q.detach().clone().requires_grad_(True) for q in params]
    ropt = torch.optim.SGD(ref_params, lr=8e-3, momentum=0.9)
    x, y = synth_images(790, 4, 30, 40)
    for _ in range(2):
        pred, acc, loss = m.observe({"image": x, "label": y})
        opt.zero_grad()
        loss.backward()
        for rp, q in zip(ref_params, params):
            rp.grad = q.grad.detach().clone()
        opt.step(); ropt.step()
    torch.cuda.synchronize()
    for rp, q in zip(ref_params, params):
        assert rel_l2(q.detach(), rp.detach()) < 1e-6
    pred, acc = m.inference({"image": x, "label": y})
    assert pred.shape == (4,) and 0.0 <= acc <= 1.0


def test_sdlora_graphed_step_equals_eager():
    from libcontinual_b200 import optim
    from libcontinual_b200.trainer import GraphedFlatStep
    p = synth_vit_state(5150)[0]
    m = _model(p)
    states = sdlora_task_states(1)
    for task in (0, 1):
        blocks, mags, hw, hb = states[task]
        m.before_task(task, None, None, None)
        _install(m, task, blocks, mags, hw, hb)
        if task == 0:
            m.after_task(0, None, None, None)
    x, y = synth_images(781, 4, 10, 20)
    theta0 = m.theta.clone()
    opt = optim.FlatSGD(m.get_parameters(None), lr=8e-3, momentum=0.9, model=m)
    for _ in range(2):
        pred, acc, loss = m.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    eager, eager_loss = m.theta.clone(), float(loss.detach())
    m.theta.copy_(theta0)
    opt2 = optim.FlatSGD(m.get_parameters(None), lr=8e-3, momentum=0.9, model=m)
    gs = GraphedFlatStep(m, opt2, 4)
    for _ in range(2):
        gs.run(x, y)
    torch.cuda.synchronize()
    assert rel_l2(m.theta, eager) < 1e-6 and abs(float(gs.loss()) - eager_loss) < 1e-5
