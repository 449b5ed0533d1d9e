"""CPU: the input-transform oracle (oracle/data_port.py) against the golden vectors written by the real torchvision / PIL
(oracle/make_golden_data.py), and the host-side draw generators of libcontinual_b200/data.py (ranges / determinism)."""
import os

import numpy as np

from oracle import data_port as dp

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "data_transforms.npz"))


def test_cifar_transform_oracle_matches_torchvision_golden():
    for i in range(G["cifar_img"].shape[0]):
        dx, dy, flip, _ = (int(v) for v in G["cifar_draw"][i])
        got = dp.cifar_transform(G["cifar_img"][i], dx, dy, bool(flip), float(G["cifar_bright"][i]))
        assert np.array_equal(got, G["cifar_out"][i]), i


def test_resize_oracle_matches_pil_golden():
    for tag in ("small", "large"):
        for i in range(G[f"{tag}_img"].shape[0]):
            t, l, h, w, oh, ow, oy, ox = (int(v) for v in G[f"{tag}_draw"][i])
            got = dp.resized_crop_window(G[f"{tag}_img"][i], t, l, h, w, oh, ow, oy, ox, 224, bool(G[f"{tag}_flip"][i]))
            want = (G[f"{tag}_out_u8"][i].astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1)
            assert np.array_equal(got, want), (tag, i)


def test_draw_generators():
    from libcontinual_b200 import data as D
    rng = np.random.default_rng(3)
    draw, bright = D.draw_cifar_train(rng, 4096)
    assert draw[:, :2].min() == 0 and draw[:, :2].max() == 8 and set(np.unique(draw[:, 2])) == {0, 1}
    assert bright.min() >= np.float32(1 - 63 / 255) and bright.max() <= np.float32(1 + 63 / 255) and 0.45 < draw[:, 2].mean() < 0.55
    d2, f2 = D.draw_resized_crop(np.random.default_rng(5), 512, 32, 32)
    d3, f3 = D.draw_resized_crop(np.random.default_rng(5), 512, 32, 32)
    assert np.array_equal(d2, d3) and np.array_equal(f2, f3)                     # seeded: reproducible
    t, l, h, w = d2[:, 0], d2[:, 1], d2[:, 2], d2[:, 3]
    assert (h >= 1).all() and (w >= 1).all() and (t + h <= 32).all() and (l + w <= 32).all() and (t >= 0).all() and (l >= 0).all()
    area = h * w / 1024.0
    assert area.min() >= 0.05 and area.max() <= 1.0 and (d2[:, 4:6] == 224).all() and (d2[:, 6:] == 0).all()
    ratio = w / h
    assert ratio.min() > 0.6 and ratio.max() < 1.6                               # (3/4, 4/3) up to integer rounding of small boxes
    d4, _ = D.draw_resize_center(1, 300, 400, 256, 224)
    assert tuple(d4[0]) == (0, 0, 300, 400, 256, 341, 16, 58)                    # Resize(256): (256, int(256*400/300)); CenterCrop offsets round((x-224)/2)
