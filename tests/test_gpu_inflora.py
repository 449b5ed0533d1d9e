"""GPU: InfLoRA_OPT on ViT-B/16 (adapter merge, rank-form adapter gradients, input-matrix pass, plugin step) through the C ABI against
(a) tests/golden/inflora_vit.npz written by the REAL reference (`core.model.InfLoRA_opt.InfLoRA_OPT.observe` + backward, `update_input_matrix`,
`merge_weight`) and (b) the oracle.  Tolerances: BF16 GEMM operands through 12 blocks forward and backward -> 3e-2 relative L2 on gradients,
2e-2 on logits / loss; the fp32 kernels (merge, rank-form gradient) to 1e-5."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.golden_util import load, synth_images, synth_lora_state, synth_vit_state
from tests.test_gpu_kernels import P, dev, lib, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


def test_lora_merge_kernel(lib):
    g = torch.Generator().manual_seed(3)
    L, D, r = 2, 768, 10
    W = torch.randn(L, 3 * D, D, generator=g) * 0.02
    A = torch.randn(L, 2, r, D, generator=g) * 0.03
    B = torch.randn(L, 2, D, r, generator=g) * 0.05
    Wd, Ad, Bd = dev(W), dev(A), dev(B)
    wb = torch.zeros(L, 3 * D, D, dtype=torch.bfloat16, device="cuda"); wbt = torch.zeros(L, D, 3 * D, dtype=torch.bfloat16, device="cuda")
    wout = torch.zeros(L, 3 * D, D, device="cuda")
    assert lib.lc_lora_merge(P(Wd), P(Ad), P(Bd), None, 0b110, L, D, r, P(wb), P(wbt), P(wout), st()) == 0
    torch.cuda.synchronize()
    ref = W.clone()
    for l in range(L):
        ref[l] = port.lora_merge_qkv(W[l], A[l, 0], B[l, 0], A[l, 1], B[l, 1])
    assert (wout.cpu()[:, D:] - ref[:, D:]).abs().max().item() < 1e-6
    assert torch.equal(wb.cpu()[:, D:], wout.cpu()[:, D:].bfloat16())
    assert torch.equal(wbt.cpu()[:, :, D:], wout.cpu()[:, D:].bfloat16().transpose(1, 2))
    assert wb[:, :D].abs().max().item() == 0 and wout[:, :D].abs().max().item() == 0          # the q slab is not touched
    # scaled, single slab (q) variant: W_q + B diag(s) A
    s = torch.rand(L, 1, r, generator=g)
    wout.zero_()
    assert lib.lc_lora_merge(P(Wd), P(dev(A[:, :1].contiguous())), P(dev(B[:, :1].contiguous())), P(dev(s)), 0b001, L, D, r, None, None, P(wout), st()) == 0
    torch.cuda.synchronize()
    refq = W[:, :D] + (B[:, 0] * s) @ A[:, 0]
    assert (wout.cpu()[:, :D] - refq).abs().max().item() < 1e-6


@pytest.mark.parametrize("n,r", [(700, 10), (25216, 10), (333, 16), (64, 4)])
def test_lora_bgrad_rows_kernel(lib, n, r):
    g = torch.Generator().manual_seed(n + r)
    D = 768
    X = (torch.randn(n, 3 * D, generator=g) * 0.1).bfloat16()
    Z = torch.randn(n, 2 * r, generator=g)
    nchunk = 37
    partial = torch.zeros(lib.lc_lora_bgrad_partial_floats(2, D, r, nchunk), device="cuda")
    out = torch.full((2, D, r), float("nan"), device="cuda")
    assert lib.lc_lora_bgrad_rows(P(dev(X)), 3 * D, D, D, 2, D, P(dev(Z)), 2 * r, r, n, P(partial), nchunk, P(out), st()) == 0
    torch.cuda.synchronize()
    Xd = X.double()
    ref = torch.stack([Xd[:, D:2 * D].T @ Z[:, :r].double(), Xd[:, 2 * D:].T @ Z[:, r:].double()])
    assert rel_l2(out, ref) < 1e-5
    # non-adjacent slabs (q, v): slab stride 2 D
    assert lib.lc_lora_bgrad_rows(P(dev(X)), 3 * D, 0, 2 * D, 2, D, P(dev(Z)), 2 * r, r, n, P(partial), nchunk, P(out), st()) == 0
    torch.cuda.synchronize()
    ref = torch.stack([Xd[:, :D].T @ Z[:, :r].double(), Xd[:, 2 * D:].T @ Z[:, r:].double()])
    assert rel_l2(out, ref) < 1e-5


def test_transpose_bf16(lib):
    g = torch.Generator().manual_seed(1)
    X = torch.randn(591, 768, generator=g).bfloat16()
    out = torch.full((768, 592), 7.0, dtype=torch.bfloat16, device="cuda")
    assert lib.lc_transpose_bf16(P(dev(X)), 768, 591, 768, P(out), 592, st()) == 0
    torch.cuda.synchronize()
    assert torch.equal(out.cpu()[:, :591], X.T) and out[:, 591].abs().max().item() == 0


def _model(p):
    from libcontinual_b200.model import InfLoRA_OPT, vit_pt_imnet
    import os
    os.environ["PYTHONHASHSEED"] = "42"
    bb = vit_pt_imnet(pretrained=False, state=p, device="cuda:0", attn_layer="MultiHeadAttention_LoRA", lora_rank=10)
    return InfLoRA_OPT(bb, "cuda:0", init_cls_num=20, inc_cls_num=20, task_num=10, lame=1.0, lamb=0.95, embd_dim=768, use_ca=False, dataset="imagenet-r")


def _install(m, task, lora, hw, hb):
    A = torch.stack([torch.stack([d["A_k"], d["A_v"]]) for d in lora])
    m.start_task(task, A)
    with torch.no_grad():
        m.lora_B.copy_(torch.stack([torch.stack([d["B_k"], d["B_v"]]) for d in lora]).cuda())
        head = m._network.classifier_pool[task]
        head.weight.copy_(hw.cuda()); head.bias.copy_(hb.cuda())


def test_inflora_observe_matches_reference_golden():
    g = load("inflora_vit.npz")
    p = synth_vit_state(5150)[0]
    m = _model(p)
    for task in (0, 1):
        lora, hw, hb = synth_lora_state(880 + task)
        _install(m, task, lora, hw, hb)
        lo = 0 if task == 0 else 20
        x, y = synth_images(700 + task, 4, lo, lo + 20)
        pred, acc, loss = m.observe({"image": x, "label": y})
        for q in m.get_parameters(None):
            q.grad = None                                        # the reference Trainer order: observe -> zero_grad -> backward (trainer.py:601-604)
        loss.backward()
        torch.cuda.synchronize()
        assert not m.engine.tensor_core_error()
        assert abs(float(loss) - float(g[f"t{task}/loss"])) < 2e-2 * abs(float(g[f"t{task}/loss"]))
        head = m._network.classifier_pool[task]
        got = {"dW": head.weight.grad, "db": head.bias.grad, "dB_k": torch.stack([q.grad for q in m.lora_B_k]), "dB_v": torch.stack([q.grad for q in m.lora_B_v])}
        for k, v in got.items():
            e = rel_l2(v, torch.from_numpy(g[f"t{task}/{k}"]))
            print(f"task{task} {k}: rel-L2 vs reference = {e:.2e}")
            assert e < 3e-2, (task, k, e)
        per_layer = [rel_l2(m.lora_B_k[i].grad, torch.from_numpy(g[f"t{task}/dB_k"][i])) for i in range(12)]
        print("dB_k per block:", ["%.1e" % e for e in per_layer])
        assert max(per_layer) < 6e-2
        bufs = m._batch_bufs(4)
        assert rel_l2(bufs["logits"][:, lo:lo + 20], torch.from_numpy(g[f"t{task}/logits"])) < 2e-2
        assert pred.shape == (4,) and int(pred.min()) >= 0 and int(pred.max()) < 20
        if task == 0:
            # the input-matrix pass with the adapters applied, and the task-0 basis
            xs = [synth_images(710 + j, 3, 0, 20)[0] for j in range(2)]
            cur = m.input_matrices(xs)
            torch.cuda.synchronize()
            proj = torch.from_numpy(np.random.default_rng(99).standard_normal((768, 8)).astype(np.float32)).cuda()
            e = rel_l2(cur @ proj, torch.from_numpy(g["cov/proj"]))
            print("input matrices (projected) rel-L2:", e)
            assert e < 1e-2
            tr = torch.stack([c.trace() for c in cur]).cpu().numpy()
            assert np.allclose(tr, g["cov/trace"], rtol=5e-3)
            for i in (0, 11):
                U, S, _ = torch.linalg.svd(cur[i], full_matrices=False)
                assert np.allclose(S[:16].cpu().numpy(), g[f"cov/S{i}"], rtol=2e-2)
            m.engine.lora_merge(w_out=True); m.engine.lora.active = False          # merge_weight() (after_task without the loader pass)
            ref_merged = port.lora_merge_qkv(p["transformer.blocks.3.attn.qkv.weight"], lora[3]["A_k"], lora[3]["B_k"], lora[3]["A_v"], lora[3]["B_v"])
            assert (m.engine.qkv_w[3].cpu() - ref_merged).abs().max().item() < 1e-6


def test_inflora_trainer_order_flat_sgd_and_graph():
    """observe -> zero_grad -> backward -> FlatSGD.step equals torch.optim.SGD on the same gradients; the CUDA-graph step reproduces the eager one."""
    from libcontinual_b200 import optim
    from libcontinual_b200.trainer import GraphedFlatStep
    p = synth_vit_state(5150)[0]
    m = _model(p)
    lora, hw, hb = synth_lora_state(880)
    _install(m, 0, lora, hw, hb)
    params = m.get_parameters(None)
    opt = optim.FlatSGD(params, lr=8e-3, momentum=0.9, model=m)
    ref_params = [q.detach().clone().requires_grad_(True) for q in params]
    ropt = torch.optim.SGD(ref_params, lr=8e-3, momentum=0.9)
    x, y = synth_images(700, 4, 0, 20)
    theta0 = m.theta.clone()
    losses = []
    for _ in range(2):
        pred, acc, loss = m.observe({"image": x, "label": y})
        opt.zero_grad()
        loss.backward()
        for rp, q in zip(ref_params, params):
            rp.grad = q.grad.detach().clone()
        opt.step(); ropt.step()
        losses.append(float(loss))
    torch.cuda.synchronize()
    for rp, q in zip(ref_params, params):
        assert rel_l2(q.detach(), rp.detach()) < 1e-6
    theta_eager = m.theta.clone()
    assert losses[1] < losses[0]
    # untouched: heads of the other tasks
    lo = m.oW + 20 * 768
    assert torch.equal(m.theta[lo:m.ob], theta0[lo:m.ob])
    # graphed
    m.theta.copy_(theta0)
    opt2 = optim.FlatSGD(params, lr=8e-3, momentum=0.9, model=m)
    gs = GraphedFlatStep(m, opt2, 4)
    for _ in range(2):
        gs.run(x, y)
    torch.cuda.synchronize()
    assert rel_l2(m.theta, theta_eager) < 1e-6
    assert abs(float(gs.loss()) - losses[1]) < 1e-5
    pred, acc = m.inference({"image": x, "label": y})
    assert pred.shape == (4,) and 0.0 <= acc <= 1.0


def test_inflora_opt_task_boundary_flow():
    """The whole plugin flow with (tiny) loaders: before_task (input-matrix pass + SVD -> lora_A), a training step, after_task (merge_weight + DualGPM
    update), before_task of the next task (basis outside the kept subspace), inference."""
    from libcontinual_b200 import optim
    p = synth_vit_state(5150)[0]
    m = _model(p)
    loader0 = [{"image": synth_images(900 + j, 4, 0, 20)[0], "label": synth_images(900 + j, 4, 0, 20)[1]} for j in range(2)]
    m.before_task(0, None, loader0, None)
    A0 = m.engine.lora.A.clone()                                   # [L, 2, r, 768]
    gram = A0[3, 0] @ A0[3, 0].T * 3.0
    assert torch.allclose(gram, torch.eye(10, device="cuda"), atol=1e-3)          # rows = orthonormal basis / sqrt(3)
    assert torch.equal(A0[:, 0], A0[:, 1]) and float(m.lora_B.abs().max()) == 0.0
    opt = optim.FlatSGD(m.get_parameters(None), lr=8e-3, momentum=0.9, model=m)
    w0 = m.engine.qkv_w.clone()
    for b in loader0:
        pred, acc, loss = m.observe(b)
        opt.zero_grad(); loss.backward(); opt.step()
    assert float(m.lora_B.abs().max()) > 0.0
    B_trained, A_used = m.lora_B.clone(), m.engine.lora.A.clone()
    m.after_task(0, None, loader0, None)
    # merge_weight: k / v slabs moved by B A, q untouched
    delta = m.engine.qkv_w - w0
    assert float(delta[:, :768].abs().max()) == 0.0
    want = B_trained[5, 0] @ A_used[5, 0]
    assert float((delta[5, 768:1536] - want).abs().max()) < 1e-6
    assert len(m.feature_list) == 12 and all(t == "remove" for t in m.project_type) and all(f.shape[1] >= 1 for f in m.feature_list)
    loader1 = [{"image": synth_images(910 + j, 4, 20, 40)[0], "label": synth_images(910 + j, 4, 20, 40)[1]} for j in range(2)]
    m.before_task(1, None, loader1, None)
    assert m._known_classes == 20 and float(m.lora_B.abs().max()) == 0.0
    # the new basis lies outside the subspace kept from task 0 ('remove' projection)
    F3 = torch.from_numpy(np.ascontiguousarray(m.feature_list[3])).cuda()
    A1 = m.engine.lora.A[3, 0]
    assert float((A1 @ F3).abs().max()) < 5e-3
    pred, acc, loss = m.observe(loader1[0])
    assert torch.isfinite(loss.detach()).all() and pred.shape == (4,)
    ipred, iacc = m.inference(loader1[0])
    assert int(ipred.max()) < 40 and not m.engine.tensor_core_error()


def test_graphed_step_prefetch_equals_plain_run():
    """`prefetch(next batch)` + `run(batch)` (double-buffered H2D on a copy stream) gives the same parameters as `run(batch)` alone."""
    from libcontinual_b200 import optim
    from libcontinual_b200.trainer import GraphedFlatStep
    p = synth_vit_state(5150)[0]
    m = _model(p)
    lora, hw, hb = synth_lora_state(880)
    _install(m, 0, lora, hw, hb)
    batches = [tuple(t.pin_memory() for t in synth_images(940 + j, 4, 0, 20)) for j in range(3)]
    theta0 = m.theta.clone()
    outs = []
    for use_prefetch in (False, True):
        m.theta.copy_(theta0)
        opt = optim.FlatSGD(m.get_parameters(None), lr=8e-3, momentum=0.9, model=m)
        gs = GraphedFlatStep(m, opt, 4)
        if use_prefetch:
            gs.prefetch(*batches[0])
        for j in range(3):
            gs.run(*batches[j])
            if use_prefetch and j + 1 < 3:
                gs.prefetch(*batches[j + 1])
        torch.cuda.synchronize()
        outs.append(m.theta.clone())
    assert torch.equal(outs[0], outs[1])
