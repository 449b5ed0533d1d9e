"""GPU: the device-side input pipeline (lc_augment_cifar_u8, lc_resize_crop_u8, GpuLoader) against the transform oracle and the torchvision / PIL golden
vectors — bit-exact — and the evaluation loop (validate / lc_eval_meter / lc_eval_fold) against a restatement of Trainer._validate."""
import os

import numpy as np
import pytest
import torch

from oracle import data_port as dp

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_transforms.npz"))


def test_cifar_augment_kernel_matches_torchvision_golden_and_oracle():
    from libcontinual_b200.data import transform_batch
    img = torch.from_numpy(G["cifar_img"]).cuda()
    out = transform_batch(img, "cifar_train", G["cifar_draw"], G["cifar_bright"]).cpu().numpy()
    assert np.array_equal(out, G["cifar_out"])
    # a full-size batch of random draws against the oracle
    rng = np.random.default_rng(11)
    B = 128
    imgs = rng.integers(0, 256, (B, 32, 32, 3), dtype=np.uint8)
    from libcontinual_b200.data import draw_cifar_train, draw_identity
    draw, bright = draw_cifar_train(rng, B)
    got = transform_batch(torch.from_numpy(imgs).cuda(), "cifar_train", draw, bright).cpu().numpy()
    for b in range(B):
        want = dp.cifar_transform(imgs[b], int(draw[b, 0]), int(draw[b, 1]), bool(draw[b, 2]), float(bright[b]))
        assert np.array_equal(got[b], want), b
    d0, _ = draw_identity(B, 4)
    got = transform_batch(torch.from_numpy(imgs).cuda(), "cifar_test", d0, None).cpu().numpy()
    assert np.array_equal(got[5], dp.cifar_transform(imgs[5]))


@pytest.mark.parametrize("tag", ["small", "large"])
def test_resize_kernel_matches_pil_golden(tag):
    from libcontinual_b200.data import transform_batch
    img = torch.from_numpy(G[f"{tag}_img"]).cuda()
    out = transform_batch(img, "vit_train", G[f"{tag}_draw"], G[f"{tag}_flip"]).cpu().numpy()
    want = (G[f"{tag}_out_u8"].astype(np.float32) / np.float32(255.0)).transpose(0, 3, 1, 2)
    assert np.array_equal(out, want)


def test_resize_kernel_random_boxes_vs_oracle():
    from libcontinual_b200.data import draw_resized_crop, transform_batch
    rng = np.random.default_rng(17)
    for (H, W, B) in ((32, 32, 16), (180, 240, 4)):
        imgs = rng.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
        draw, flip = draw_resized_crop(rng, B, H, W)
        got = transform_batch(torch.from_numpy(imgs).cuda(), "vit_train", draw, flip).cpu().numpy()
        for b in range(B):
            t, l, h, w, oh, ow, oy, ox = (int(v) for v in draw[b])
            assert np.array_equal(got[b], dp.resized_crop_window(imgs[b], t, l, h, w, oh, ow, oy, ox, 224, bool(flip[b]))), (H, b)


def test_gpu_loader_batches_and_labels():
    from libcontinual_b200.data import DeviceImageDataset, GpuLoader
    rng = np.random.default_rng(23)
    N = 300
    imgs = rng.integers(0, 256, (N, 32, 32, 3), dtype=np.uint8)
    labels = rng.integers(0, 10, N)
    ds = DeviceImageDataset(imgs, labels)
    ld = GpuLoader(ds, 128, "cifar_train", shuffle=True, drop_last=False, seed=5)
    assert len(ld) == 3 and ld.dataset is ds
    seen = []
    for batch in ld:
        x, y = batch["image"], batch["label"]
        idx, draw, bright = ld._last_draw
        assert x.is_cuda and y.is_cuda and x.shape[1:] == (3, 32, 32) and x.shape[0] == len(idx) == y.shape[0]
        assert np.array_equal(y.cpu().numpy(), labels[idx])
        b = len(idx) // 2
        want = dp.cifar_transform(imgs[idx[b]], int(draw[b, 0]), int(draw[b, 1]), bool(draw[b, 2]), float(bright[b]))
        assert np.array_equal(x[b].cpu().numpy(), want)
        seen.append(idx)
    assert sorted(np.concatenate(seen).tolist()) == list(range(N))                      # one epoch = every sample once
    again = [b["image"].clone() for b in GpuLoader(ds, 128, "cifar_train", shuffle=True, seed=5)]
    first = [b["image"].clone() for b in GpuLoader(ds, 128, "cifar_train", shuffle=True, seed=5)]
    assert all(torch.equal(a, b) for a, b in zip(again, first))                         # seeded: reproducible
    # the ViT test transform through the loader: Resize(224) of every sample, no randomness
    lv = GpuLoader(ds, 64, "vit_test", shuffle=False)
    xb = next(iter(lv))["image"]
    assert xb.shape == (64, 3, 224, 224)
    assert np.array_equal(xb[3].cpu().numpy(), dp.resized_crop_window(imgs[3], 0, 0, 32, 32, 224, 224, 0, 0, 224, False))


class _FakeModel(torch.nn.Module):
    """inference() answers from a table: pred = table[label-keyed position]; counts how it was called."""

    def __init__(self, preds):
        super().__init__()
        self.w = torch.nn.Parameter(torch.zeros(1, device="cuda"))
        self.preds, self.cursor, self.calls = preds, 0, []

    def inference(self, data, task_id=None):
        from libcontinual_b200._lib import host_acc
        n = data["label"].shape[0]
        p = self.preds[self.cursor:self.cursor + n].cuda()
        self.cursor += n
        self.calls.append(task_id)
        return p, host_acc(self, (p == data["label"].cuda()).sum(), n)


def _reference_validate(preds, loaders, per_task, bounds=None):
    """core/trainer.py:616-720 restated on the host."""
    cur, per, ca, na = 0, [], 0, 0
    if per_task:
        for ld in loaders:
            c, n = 0, 0
            for b in ld:
                m = b["label"].shape[0]
                acc = float((preds[cur:cur + m] == b["label"]).sum().item()) / m
                cur += m
                c += int(acc * m); n += m
            ca += c; na += n
            per.append(round(c * 100 / n, 2))
    else:
        cb, nb = np.zeros(len(bounds) - 1, dtype=int), np.zeros(len(bounds) - 1, dtype=int)
        for ld in loaders:
            for b in ld:
                m = b["label"].shape[0]
                p, y = preds[cur:cur + m].numpy(), b["label"].numpy()
                cur += m
                ca += int((p == y).sum()); na += m
                for t in range(len(bounds) - 1):
                    mk = (y >= bounds[t]) & (y < bounds[t + 1])
                    cb[t] += int((p[mk] == y[mk]).sum()); nb[t] += int(mk.sum())
        per = [round(c * 100 / n, 2) if n > 0 else 0 for c, n in zip(cb, nb)]
    return {"avg_acc": round(ca * 100 / na, 2), "per_task_acc": per}


@pytest.mark.parametrize("per_task", [True, False])
def test_validate_matches_reference_loop(per_task):
    from libcontinual_b200.trainer import validate
    g = torch.Generator().manual_seed(9)
    loaders, labels = [], []
    for t in range(3):
        lo, hi = (0, 10) if t == 0 else (10 + 5 * (t - 1), 15 + 5 * (t - 1))
        bs = []
        for m in (100, 100, 37):                       # batch sizes where int(acc * m) loses a sample for some counts (e.g. 29 / 100)
            y = torch.randint(lo, hi, (m,), generator=g)
            bs.append({"image": torch.zeros(m, 3, 4, 4), "label": y})
            labels.append(y)
        loaders.append(bs)
    y_all = torch.cat(labels)
    preds = torch.where(torch.rand(y_all.shape, generator=g) < 0.29, y_all, (y_all + 1) % 20)
    preds[:100] = (y_all[:100] + 1) % 20
    preds[:29] = y_all[:29]                            # first batch: exactly 29 / 100 correct -> the reference counts 28
    want = _reference_validate(preds, loaders, per_task, bounds=[0, 10, 15, 20])
    model = _FakeModel(preds)
    got = validate(model, loaders, 2, setting="task-agnostic", testing_per_task=per_task, init_cls_num=10, inc_cls_num=5)
    assert got == want, (got, want)
    assert not getattr(model, "_defer_metrics", False)
    if per_task:
        assert int(0.29 * 100) == 28                   # the quirk the per-task mode reproduces
    model2 = _FakeModel(preds)
    validate(model2, loaders, 2, setting="task-aware", testing_per_task=True)
    assert model2.calls == [0] * 3 + [1] * 3 + [2] * 3
