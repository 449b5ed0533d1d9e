"""Helpers shared by the oracle-vs-golden (CPU) and CUDA-vs-oracle (GPU) tests.  Inputs and initial weights are
regenerated from numpy PCG64 seeds exactly as `oracle/make_golden.py` did; only reference OUTPUTS live in the
.npz fixtures."""
import os

import numpy as np
import torch

from oracle import port

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


def synth_resnet_state(seed, n_head, feat=64):
    rng = np.random.default_rng(seed)
    p, b = port.cifar_resnet_init(rng)
    bound = 1.0 / np.sqrt(feat)
    fc_w = torch.from_numpy(rng.uniform(-bound, bound, (n_head, feat)).astype(np.float32))
    fc_b = torch.from_numpy(rng.uniform(-bound, bound, (n_head,)).astype(np.float32))
    return p, b, fc_w, fc_b


def synth_batch(seed, B, lo, hi, img=32):
    rng = np.random.default_rng(seed)
    x = torch.from_numpy(rng.standard_normal((B, 3, img, img)).astype(np.float32))
    y = torch.from_numpy(rng.integers(lo, hi, (B,)).astype(np.int64))
    return x, y


def check_summary(g, prefix, named, rtol=1e-4, atol=1e-6, full_rtol=None):
    """Compare a dict of tensors with a fixture summary written by make_golden.summarize()."""
    names = [str(n) for n in g[prefix + "/names"]]
    norms = g[prefix + "/norm"]
    for i, n in enumerate(names):
        t = named[n].detach().double().cpu()
        assert abs(float(t.norm()) - norms[i]) <= rtol * norms[i] + atol, (prefix, n, float(t.norm()), norms[i])
        key = prefix + "/full/" + n
        if key in g.files:
            ref = torch.from_numpy(g[key]).double()
            scale = float(ref.abs().max()) + 1e-12
            err = float((t - ref).abs().max()) / scale
            assert err <= (full_rtol or rtol) , (prefix, n, err)
