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
