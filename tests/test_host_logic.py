"""CPU: host-side task-boundary logic of the method plugins against outputs of the REAL reference (fixtures written by oracle/make_golden.py)."""
import numpy as np

from tests.golden_util import load, synth_input_matrices


def test_dualgpm_update_matches_reference():
    """`dualgpm_update` (InfLoRA_OPT.after_task bookkeeping) vs `InfLoRA_OPT._update_feature` (InfLoRA_opt.py:278-362) on preset input matrices:
    basis sizes, 'remove' / 'retain' types (the swap happens at task 2 for blocks 10, 11) and projectors, for three blocks over three tasks.
    The 'retain' shrink branch is not pinned: the reference itself raises there under numpy 2 (fixture key `unpinned_from_task`)."""
    from libcontinual_b200.model.inflora import dualgpm_update
    g = load("dualgpm.npz")
    layers = [0, 10, 11]
    proj = np.random.default_rng(7).standard_normal((768, 4)).astype(np.float32)
    feats, types = [], []
    for task in range(3):
        acts = synth_input_matrices(1200 + task)[layers]
        dualgpm_update(acts, feats, types, task, 4, 0.9999, 0.999)
        assert [f.shape[1] for f in feats] == g[f"t{task}/sizes"][layers].tolist()
        assert [t == "retain" for t in types] == g[f"t{task}/types"][layers].tolist()
        P = np.stack([f @ (f.T @ proj) for f in feats])
        assert np.allclose(P, g[f"t{task}/P"][layers], rtol=1e-3, atol=1e-3)
    assert types[1] == "retain" and types[0] == "remove"
    assert int(g["unpinned_from_task"]) == 3


def test_dualgpm_update_v1_matches_reference():
    """`dualgpm_update_v1` vs `InfLoRA.update_DualGPM` (InfLoRA.py:215-307) over four sessions: growth, the 'remove' -> 'retain' swap (session 2) and the
    'retain' shrink (session 3), sizes / types exact, projectors to round-off."""
    from libcontinual_b200.model.inflora_orig import dualgpm_update_v1
    g = load("inflora_orig_vit.npz")
    assert int(g["gpm/unpinned_from_task"]) == 4
    proj = np.random.default_rng(7).standard_normal((768, 4)).astype(np.float32)
    feats, types = [], []
    for task in range(4):
        acts = synth_input_matrices(1200 + task)[[0, 10, 11]]
        dualgpm_update_v1(list(acts), feats, types, task, 4, 0.9999, 0.999)
        assert [f.shape[1] for f in feats] == g[f"gpm/t{task}/sizes"].tolist()
        assert [t == "retain" for t in types] == g[f"gpm/t{task}/types"].tolist()
        P = np.stack([f @ (f.T @ proj) for f in feats])
        assert np.allclose(P, g[f"gpm/t{task}/P"], rtol=1e-3, atol=1e-3)
    assert types == ["remove", "retain", "retain"]
