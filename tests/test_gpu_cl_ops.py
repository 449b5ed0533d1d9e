"""GPU: continual-learning specific kernels through the C ABI vs the CPU oracle (oracle/port.py) and the golden vectors recorded
from the reference (tests/golden/ops_small.npz).  Integer outputs (prompt ids, histograms, predictions) exact; fp32 rtol 1e-4."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import port
from tests.golden_util import load
from tests.test_gpu_kernels import P, close, dev, lib, st, _keepalive  # noqa: F401

pytestmark = pytest.mark.gpu


def test_l2p_select_matches_reference_golden_and_oracle(lib):
    g = load("ops_small.npz")
    rng = np.random.default_rng(404)
    for case in range(4):
        B, pool, topk, length, D = [int(v) for v in g[f"l2p{case}/shape"]]
        prm = torch.from_numpy(rng.uniform(0, 1, (1, pool, length, D)).astype(np.float32))
        key = torch.from_numpy(rng.uniform(0, 1, (pool, D)).astype(np.float32))
        q = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32))
        if case == 3:
            key[1] = key[0]; key[4] = key[0]
        sim = torch.zeros(B, pool, device="cuda"); ids = torch.zeros(topk, dtype=torch.int64, device="cuda")
        hist = torch.zeros(pool, dtype=torch.int32, device="cuda"); rs = torch.zeros(1, device="cuda")
        dkey = torch.zeros(pool, D, device="cuda"); scratch = torch.zeros(D, device="cuda")
        assert lib.lc_l2p_select(P(dev(q)), P(dev(key)), B, pool, D, topk, P(sim), P(ids), P(hist), P(rs), P(dkey), P(scratch), st()) == 0
        kn, qn = F.normalize(key, dim=-1), F.normalize(q, dim=-1)
        close(sim, qn @ kn.T, 1e-5, 1e-6, "similarity")
        ids_rule = port.l2p_majority_ids_numpy((qn @ kn.T).numpy(), topk)
        if case != 3:          # exact ties in the similarity (case 3) make the fp32 top-k itself order dependent
            assert np.array_equal(ids.cpu().numpy(), ids_rule), (case, ids.cpu().numpy(), ids_rule)
            assert np.array_equal(hist.cpu().numpy(), g[f"l2p{case}/hist"])
            if int(g[f"l2p{case}/strict"]):
                assert np.array_equal(ids.cpu().numpy(), g[f"l2p{case}/ids"])          # the reference's own ids
                assert abs(float(rs) - float(g[f"l2p{case}/reduce_sim"])) < 1e-5
                close(dkey, -torch.from_numpy(g[f"l2p{case}/dkey"]), 1e-4, 1e-6, "d(reduce_sim)/d(key)")   # fixture holds d(-rs)
        # gather
        idsel = ids.clone()
        out = torch.empty(B, topk * length, D, device="cuda")
        assert lib.lc_l2p_gather(P(dev(prm[0])), P(idsel), P(out), B, topk, length, D, st()) == 0
        ref = prm[0][idsel.cpu()].reshape(topk * length, D).unsqueeze(0).expand(B, -1, -1)
        close(out, ref, 0, 0, "gathered prompts")


def test_cosine_head_and_lucir_loss_vs_oracle(lib):
    g = load("ops_small.npz")
    rng = np.random.default_rng(11)
    B, D, n_old, n_new, K = 32, 64, 10, 5, 2
    C = n_old + n_new
    feat = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32)).requires_grad_(True)
    ref_feat = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32))
    W = torch.from_numpy(rng.uniform(-0.125, 0.125, (C, D)).astype(np.float32)).requires_grad_(True)
    sigma = torch.tensor([2.5], requires_grad=True)
    y = torch.from_numpy(rng.integers(0, C, (B,)).astype(np.int64))
    cur_lamda, margin, lw_mr = 5.0 * np.sqrt(n_old / n_new), 0.5, 1.0
    scores_ref = port.cosine_head(feat, W, None)
    logits_ref = sigma * scores_ref
    loss_ref = port.lucir_loss(feat, ref_feat, logits_ref, scores_ref, y, n_old, cur_lamda, K, margin, lw_mr)
    dfeat_ref, dW_ref, dsig_ref = torch.autograd.grad(loss_ref, [feat, W, sigma])
    # forward
    ld = 100
    inv = torch.zeros(B + C, device="cuda"); scores = torch.zeros(B, ld, device="cuda"); logits = torch.zeros(B, ld, device="cuda")
    fd, Wd, sd = dev(feat.detach()), dev(W.detach()), dev(sigma.detach())
    assert lib.lc_cosine_head_forward(P(fd), P(Wd), P(sd), B, C, D, P(inv), P(scores), P(logits), ld, st()) == 0
    close(scores[:, :C], scores_ref, 1e-5, 1e-6, "cosine scores"); close(logits[:, :C], logits_ref, 1e-5, 1e-6, "cosine logits")
    # the reference's own CosineLinear output on its fixture inputs
    rng2 = np.random.default_rng(404)
    for case in range(4):      # replay the draws that precede the cosine fixtures
        Bq, pool, topk, length, Dq = [int(v) for v in g[f"l2p{case}/shape"]]
        rng2.uniform(0, 1, (1, pool, length, Dq)); rng2.uniform(0, 1, (pool, Dq)); rng2.standard_normal((Bq, Dq))
    f2 = torch.from_numpy(rng2.standard_normal((16, 64)).astype(np.float32))
    w1 = torch.from_numpy(rng2.uniform(-0.125, 0.125, (10, 64)).astype(np.float32))
    w2 = torch.from_numpy(rng2.uniform(-0.125, 0.125, (5, 64)).astype(np.float32))
    inv2 = torch.zeros(31, device="cuda"); s2 = torch.zeros(16, ld, device="cuda"); l2 = torch.zeros(16, ld, device="cuda")
    assert lib.lc_cosine_head_forward(P(dev(f2)), P(dev(torch.cat([w1, w2]))), P(dev(torch.tensor([2.5]))), 16, 15, 64, P(inv2), P(s2), P(l2), ld, st()) == 0
    close(l2[:, :15], torch.from_numpy(g["cos/split_out"]), 1e-5, 1e-6, "SplitCosineLinear golden")
    # loss
    dl = torch.zeros(B, ld, device="cuda"); ds = torch.zeros(B, ld, device="cuda"); dfe = torch.zeros(B, D, device="cuda")
    pred = torch.zeros(B, dtype=torch.int64, device="cuda"); scal = torch.zeros(8, device="cuda"); dsg = torch.zeros(1, device="cuda")
    assert lib.lc_lucir_loss(P(logits), P(scores), ld, P(fd), P(dev(ref_feat)), D, P(dev(y)), B, C, n_old, K, float(cur_lamda), margin, lw_mr,
                             P(dl), P(ds), P(dfe), P(pred), P(scal), P(dsg), st()) == 0
    assert abs(float(scal[0]) - float(loss_ref)) <= 1e-5 * abs(float(loss_ref)) + 1e-6, (float(scal[0]), float(loss_ref))
    assert torch.equal(pred.cpu(), logits_ref.argmax(1))
    # backward through the head: gscores = sigma*dlogits + dscores ; dsigma = sum(dlogits * scores)
    gs = (2.5 * dl + ds).contiguous()
    dfeat = torch.zeros(B, D, device="cuda"); dW = torch.zeros(C, D, device="cuda")
    assert lib.lc_cosine_head_backward(P(gs), ld, P(fd), P(Wd), P(inv), B, C, D, P(dfeat), P(dW), st()) == 0
    close(dfeat + dfe, dfeat_ref, 1e-4, 1e-6, "d feat"); close(dW, dW_ref, 1e-4, 1e-6, "d W")
    assert abs(float((dl[:, :C] * scores[:, :C]).sum()) - float(dsig_ref)) < 1e-5 and abs(float(dsg) - float(dsig_ref)) < 1e-5


def test_gpm_project_and_lora(lib):
    rng = np.random.default_rng(3)
    for R, D in [(64, 48), (128, 576), (256, 512), (37, 2048)]:
        gmat = torch.from_numpy(rng.standard_normal((R, D)).astype(np.float32))
        U = torch.linalg.qr(torch.from_numpy(rng.standard_normal((D, max(4, D // 10))).astype(np.float32)))[0]
        M = (U @ U.T).contiguous()
        gd = dev(gmat.clone())
        assert lib.lc_gpm_project(P(gd), P(dev(M)), R, D, st()) == 0
        ref = port.gpm_project(gmat, M)
        close(gd, ref, 1e-4, 1e-4, f"gpm {R}x{D}")
        assert float((gd.cpu() @ U).abs().max()) < 1e-3          # projected gradient is orthogonal to the stored basis
    D, r = 768, 10
    qkv = torch.from_numpy(rng.standard_normal((3 * D, D)).astype(np.float32) * 0.02)
    Ak, Av = (torch.from_numpy(rng.standard_normal((r, D)).astype(np.float32) * 0.1) for _ in range(2))
    Bk, Bv = (torch.from_numpy(rng.standard_normal((D, r)).astype(np.float32) * 0.1) for _ in range(2))
    out = torch.empty(3 * D, D, device="cuda")
    assert lib.lc_lora_merge_qkv(P(dev(qkv)), P(dev(Ak)), P(dev(Bk)), P(dev(Av)), P(dev(Bv)), P(out), D, r, st()) == 0
    close(out, port.lora_merge_qkv(qkv, Ak, Bk, Av, Bv), 1e-5, 1e-6, "lora merge")
    zero = torch.zeros(D, r)
    assert lib.lc_lora_merge_qkv(P(dev(qkv)), P(dev(Ak)), P(dev(zero)), P(dev(Av)), P(dev(zero)), P(out), D, r, st()) == 0
    assert torch.equal(out.cpu(), qkv)                            # B = 0 reproduces the frozen weights bit for bit
    dWk = torch.from_numpy(rng.standard_normal((D, D)).astype(np.float32))
    dB = torch.empty(D, r, device="cuda")
    assert lib.lc_lora_bgrad(P(dev(dWk)), P(dev(Ak)), P(dB), D, r, st()) == 0
    close(dB, dWk @ Ak.T, 1e-4, 1e-4, "lora dB")


def test_herding_and_ncm_bit_exact_indices(lib):
    """Exemplar selection and nearest-class-mean classification: integer outputs must equal the oracle's exactly."""
    rng = np.random.default_rng(21)
    for ncls, per_cls_n, pick in [(5, 500, 20), (10, 37, 40), (3, 64, 64)]:
        sizes = [per_cls_n - (c % 3) for c in range(ncls)]
        feats = torch.from_numpy(rng.standard_normal((sum(sizes), 64)).astype(np.float32)).abs()
        feats = feats / feats.norm(dim=1).view(-1, 1)
        targets = torch.cat([torch.full((s,), c, dtype=torch.int64) for c, s in enumerate(sizes)])
        ref = port.herding_select(feats, targets, pick)
        begins = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
        out = torch.zeros(ncls, pick, dtype=torch.int64, device="cuda")
        work = torch.empty_like(feats, device="cuda")
        assert lib.lc_herding_select(P(dev(feats)), P(dev(begins)), ncls, 64, pick, P(work), P(out), st()) == 0
        got = [int(v) for v in out.cpu().flatten() if v >= 0]
        assert got == ref, (ncls, per_cls_n, pick)
    feats = torch.from_numpy(rng.standard_normal((200, 64)).astype(np.float32))
    means = torch.from_numpy(rng.standard_normal((55, 64)).astype(np.float32))
    pred = torch.zeros(200, dtype=torch.int64, device="cuda")
    assert lib.lc_ncm_classify(P(dev(feats)), P(dev(means)), 200, 55, 64, P(pred), st()) == 0
    assert torch.equal(pred.cpu(), port.ncm_classify(feats, means))


def test_herding_matches_reference_golden(lib):
    from oracle.make_golden import herding_inputs
    raw, labels = herding_inputs()
    feats = raw / raw.norm(dim=1).view(-1, 1)
    sizes = np.bincount(labels.numpy())
    begins = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    out = torch.zeros(len(sizes), 25, dtype=torch.int64, device="cuda")
    work = torch.empty_like(feats, device="cuda")
    assert lib.lc_herding_select(P(dev(feats)), P(dev(begins)), len(sizes), 64, 25, P(work), P(out), st()) == 0
    assert [int(v) for v in out.cpu().flatten()] == [int(i) for i in load("herding.npz")["idx"]]


def test_gpm_project_tensor_core(lib):
    """The projection of every AlexNet_TRGP layer shape ([64,48], [128,576], [256,512], [2048,1024], [2048,2048]) on the tensor cores
    through the two-term BF16 split, against `g - g.view(out, -1) @ M` in float64 and against the reference op in fp32 (gpm.py:78-81)."""
    from libcontinual_b200.gpm import GPMProjector
    rng = np.random.default_rng(5)
    shapes = [((64, 3, 4, 4), 48), ((128, 64, 3, 3), 576), ((256, 128, 2, 2), 512), ((2048, 1024), 1024), ((2048, 2048), 2048)]
    feats = [torch.linalg.qr(torch.from_numpy(rng.standard_normal((D, max(4, D // 10))).astype(np.float32)))[0] for _, D in shapes]
    proj = GPMProjector(feats)
    for i, (shape, D) in enumerate(shapes):
        g = torch.from_numpy(rng.standard_normal(shape).astype(np.float32))
        M = feats[i] @ feats[i].T
        ref32 = port.gpm_project(g, M)
        ref64 = g.double() - (g.double().view(shape[0], -1) @ M.double()).view(shape)
        got = proj.project_(i, g.cuda().contiguous())
        torch.cuda.synchronize()
        assert int(proj.err.item()) == 0
        e64 = float((got.cpu().double() - ref64).norm() / ref64.norm())
        e32 = float((ref32.double() - ref64).norm() / ref64.norm())
        print(f"gpm tc {tuple(shape)}: rel-L2 vs float64 {e64:.1e} (the fp32 reference op itself: {e32:.1e})")
        assert e64 < 2e-5, (shape, e64)
        assert float((got.cpu().view(shape[0], -1) @ feats[i]).abs().max()) < 2e-3      # orthogonal to the stored basis


def test_l2p_select_phases_and_global_batch_vote(lib):
    """lc_l2p_select_phase: (1) counts + (2) vote == the one-call selection bit for bit; and the data-parallel use — counts of two half batches,
    histograms summed (what the SUM all-reduce does), vote on the sum — picks the ids of the WHOLE batch (prompt.py:380-401 at the global batch)."""
    rng = np.random.default_rng(909)
    B, pool, topk, D = 64, 10, 5, 768
    key = torch.from_numpy(rng.uniform(0, 1, (pool, D)).astype(np.float32))
    q = torch.from_numpy(rng.standard_normal((B, D)).astype(np.float32))

    def run(qq, phase, hist=None):
        n = qq.shape[0]
        sim = torch.zeros(n, pool, device="cuda"); ids = torch.zeros(topk, dtype=torch.int64, device="cuda")
        hist = torch.zeros(pool, dtype=torch.int32, device="cuda") if hist is None else hist
        rs = torch.zeros(1, device="cuda"); dkey = torch.zeros(pool, D, device="cuda"); scratch = torch.zeros(D, device="cuda")
        qd, kd = dev(qq), dev(key)
        if phase == "one":
            assert lib.lc_l2p_select(P(qd), P(kd), n, pool, D, topk, P(sim), P(ids), P(hist), P(rs), P(dkey), P(scratch), st()) == 0
        elif phase == "two":
            assert lib.lc_l2p_select_phase(P(qd), P(kd), n, pool, D, topk, P(sim), P(ids), P(hist), P(rs), P(dkey), P(scratch), 1, st()) == 0
            assert lib.lc_l2p_select_phase(P(qd), P(kd), n, pool, D, topk, P(sim), P(ids), P(hist), P(rs), P(dkey), P(scratch), 2, st()) == 0
        elif phase == 1:
            assert lib.lc_l2p_select_phase(P(qd), P(kd), n, pool, D, topk, P(sim), P(ids), P(hist), P(rs), P(dkey), P(scratch), 1, st()) == 0
        else:
            assert lib.lc_l2p_select_phase(P(qd), P(kd), n, pool, D, topk, P(sim), P(ids), P(hist), P(rs), P(dkey), P(scratch), 2, st()) == 0
        torch.cuda.synchronize()
        return sim.cpu(), ids.cpu(), hist, rs.cpu(), dkey.cpu()

    one = run(q, "one")
    two = run(q, "two")
    for a, b in zip(one, two):
        assert torch.equal(a.cpu(), b.cpu())
    assert lib.lc_l2p_select_phase(None, None, B, pool, D, topk, None, None, None, None, None, None, 3, st()) != 0      # bad phase / pointers: error code

    # two "ranks": counts per half, summed histogram, vote on each half with the global histogram
    h0 = run(q[:32], 1)[2]
    h1 = run(q[32:], 1)[2]
    total = (h0 + h1).clone()
    assert torch.equal(total.cpu(), one[2].cpu())                       # the summed counts ARE the whole batch's histogram
    r0 = run(q[:32], 2, hist=total.clone())
    r1 = run(q[32:], 2, hist=total.clone())
    assert torch.equal(r0[1], one[1]) and torch.equal(r1[1], one[1])   # both ranks pick the whole batch's ids
    kn, qn = F.normalize(key, dim=-1), F.normalize(q, dim=-1)
    assert np.array_equal(one[1].numpy(), port.l2p_majority_ids_numpy((qn @ kn.T).numpy(), topk))
    # the mean of the two ranks' pull-constraint terms (what DDP's loss / gradient averaging yields) is the whole batch's
    assert abs(0.5 * (float(r0[3]) + float(r1[3])) - float(one[3])) < 1e-5
    close(0.5 * (r0[4] + r1[4]), one[4], 1e-4, 1e-6, "d(reduce_sim)/d(key), averaged over the two shards")
