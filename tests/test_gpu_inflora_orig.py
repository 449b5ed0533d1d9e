"""GPU: InfLoRA (original formulation: the per-task k / v adapters stay separate and their sum enters every forward; timm-style ViT-B/16) through the
C ABI against tests/golden/inflora_orig_vit.npz, written by the REAL `core.model.InfLoRA.InfLoRA.observe` + backward on the real `ViT_lora_co`.
Tolerances as for InfLoRA_OPT: 3e-2 relative L2 on gradients, 2e-2 on logits / loss."""
import numpy as np
import pytest
import torch

from tests.golden_util import load, synth_images, synth_stacked_adapters, synth_timm_vit_state

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_inflora_orig_observe_matches_reference_golden():
    from libcontinual_b200 import optim
    from libcontinual_b200.model import InfLoRA, SiNet_vit
    from libcontinual_b200.trainer import GraphedFlatStep
    g = load("inflora_orig_vit.npz")
    _, p_timm = synth_timm_vit_state(5150)
    bb = SiNet_vit(total_sessions=10, rank=10, init_cls=10, embd_dim=768, state=p_timm, device="cuda:0")
    m = InfLoRA(bb, 768, 100, inc_cls_num=10, device="cuda:0", lame=1.0, lamb=0.95, total_sessions=10)
    blocks, hw, hb = synth_stacked_adapters(990, 2)
    for task in (0, 1):
        A = torch.stack([torch.stack([blocks[l][task]["A_k"], blocks[l][task]["A_v"]]) for l in range(12)])
        m.start_task(A)
        assert m._known_classes == 10 * task and bb.numtask == task + 1
        with torch.no_grad():
            m.B_cur.copy_(torch.stack([torch.stack([blocks[l][task]["B_k"], blocks[l][task]["B_v"]]) for l in range(12)]).cuda())
            head = bb.classifier_pool[task]
            head.weight.copy_(hw[task].cuda()); head.bias.copy_(hb[task].cuda())
        lo = 10 * task
        x, y = synth_images(800 + task, 4, lo, lo + 10)
        pred, acc, loss = m.observe({"image": x, "label": y})
        for q in m.get_parameters(None):
            q.grad = None
        loss.backward()
        torch.cuda.synchronize()
        assert not m.engine.tensor_core_error()
        assert abs(float(loss.detach()) - float(g[f"t{task}/loss"])) < 2e-2 * abs(float(g[f"t{task}/loss"]))
        got = {"dW": head.weight.grad, "db": head.bias.grad, "dB_k": torch.stack([q.grad for q in m.lora_B_k]), "dB_v": torch.stack([q.grad for q in m.lora_B_v])}
        for k, v in got.items():
            e = rel_l2(v, torch.from_numpy(g[f"t{task}/{k}"]))
            print(f"task{task} {k}: rel-L2 vs reference = {e:.2e}")
            assert e < 3e-2, (task, k, e)
        assert rel_l2(m._batch_bufs(4)["logits"][:, lo:lo + 10], torch.from_numpy(g[f"t{task}/logits"])) < 2e-2
        assert np.array_equal(pred.cpu().numpy(), g[f"t{task}/pred"])
        ipred, iacc = m.inference({"image": x, "label": y})
        torch.cuda.synchronize()
        assert rel_l2(m._batch_bufs(4)["logits"][:, :lo + 10], torch.from_numpy(g[f"t{task}/interface"])) < 2e-2
    # cur-matrix pass with both adapters applied
    xs = [synth_images(810 + j, 3, 0, 20)[0] for j in range(2)]
    cur = m.input_matrices(xs)
    torch.cuda.synchronize()
    proj = torch.from_numpy(np.random.default_rng(99).standard_normal((768, 8)).astype(np.float32)).cuda()
    assert rel_l2(cur @ proj, torch.from_numpy(g["cov/proj"])) < 1e-2
    assert np.allclose(torch.stack([c.trace() for c in cur]).cpu().numpy(), g["cov/trace"], rtol=5e-3)
    # flat SGD over the active ranges, eager == CUDA graph
    params = m.get_parameters(None)
    theta0 = m.theta.clone()
    opt = optim.FlatSGD(params, lr=8e-3, momentum=0.9, model=m)
    for _ in range(2):
        pred, acc, loss = m.observe({"image": x, "label": y})
        opt.zero_grad(); loss.backward(); opt.step()
    torch.cuda.synchronize()
    eager = m.theta.clone()
    assert not torch.equal(eager[:m.nB], theta0[:m.nB])
    assert torch.equal(eager[m.oW:m.oW + 10 * 768], theta0[m.oW:m.oW + 10 * 768])          # task 0's head is frozen
    ref_params = [q.detach().clone() for q in params]
    m.theta.copy_(theta0)
    opt2 = optim.FlatSGD(params, lr=8e-3, momentum=0.9, model=m)
    gs = GraphedFlatStep(m, opt2, 4)
    yrel = (y - 10).cuda()
    for _ in range(2):
        gs.run(x, yrel)
    torch.cuda.synchronize()
    assert rel_l2(m.theta, eager) < 1e-6


def test_inflora_orig_task_boundary_flow():
    """before_task (cur-matrix pass + SVD -> lora_A of the new adapter), training steps, after_task (update_DualGPM + projection matrices), next task."""
    from libcontinual_b200 import optim
    from libcontinual_b200.model import InfLoRA, SiNet_vit
    _, p_timm = synth_timm_vit_state(5150)
    bb = SiNet_vit(total_sessions=10, rank=10, init_cls=10, embd_dim=768, state=p_timm, device="cuda:0")
    m = InfLoRA(bb, 768, 100, inc_cls_num=10, device="cuda:0", lame=1.0, lamb=0.95, total_sessions=10)
    mk = lambda seed, lo: [{"image": synth_images(seed + j, 4, lo, lo + 10)[0], "label": synth_images(seed + j, 4, lo, lo + 10)[1]} for j in range(2)]
    l0 = mk(920, 0)
    m.before_task(0, None, l0, None)
    assert bb.numtask == 1 and m.engine.lora.R == 10
    opt = optim.FlatSGD(m.get_parameters(None), lr=8e-3, momentum=0.9, model=m)
    for b in l0:
        pred, acc, loss = m.observe(b)
        opt.zero_grad(); loss.backward(); opt.step()
    B0 = m.B_cur.clone()
    assert float(B0.abs().max()) > 0.0
    m.after_task(0, None, l0, None)
    assert len(m.feature_list) == 12 and len(m.feature_mat) == 12 and m.feature_mat[0].shape == (768, 768)
    l1 = mk(930, 10)
    m.before_task(1, None, l1, None)
    assert bb.numtask == 2 and m._known_classes == 10 and m.engine.lora.R == 20
    assert torch.equal(m.B_old[:, :, :, :10], B0) and float(m.B_cur.abs().max()) == 0.0       # task 0's adapter is frozen in the stack, the new one starts at 0
    pred, acc, loss = m.observe(l1[0])
    assert torch.isfinite(loss.detach()).all() and int(pred.max()) < 10
    ipred, iacc = m.inference(l1[0])
    assert int(ipred.max()) < 20 and not m.engine.tensor_core_error()
