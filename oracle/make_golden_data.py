"""TEST INFRASTRUCTURE (not shipped).  Golden vectors of the reference's input transforms (core/data/data.py:11-36) written by the REAL torchvision / PIL
driven with explicit draws (torchvision.transforms.functional), after asserting that oracle/data_port.py reproduces them exactly.

    python oracle/make_golden_data.py      ->  tests/golden/data_transforms.npz
"""
import os
import sys

import numpy as np
from PIL import Image
import torchvision.transforms.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import data_port as dp  # noqa: E402


def main():
    rng = np.random.default_rng(20240607)
    out = {}
    # --- CIFAR train transform: RandomCrop(32, 4) + flip + brightness + ToTensor + Normalize ----------------------------------------------------------
    n = 12
    imgs = rng.integers(0, 256, (n, 32, 32, 3), dtype=np.uint8)
    draw = np.zeros((n, 4), dtype=np.int32)
    draw[:, 0] = rng.integers(0, 9, n); draw[:, 1] = rng.integers(0, 9, n); draw[:, 2] = rng.integers(0, 2, n)
    draw[0] = (4, 4, 0, 0); draw[1] = (0, 8, 1, 0); draw[2] = (8, 0, 0, 0)
    bright = rng.uniform(1 - 63 / 255, 1 + 63 / 255, n).astype(np.float32)
    bright[0] = 1.0; bright[1] = np.float32(1 + 63 / 255); bright[2] = np.float32(1 - 63 / 255)
    ref = np.zeros((n, 3, 32, 32), dtype=np.float32)
    for i in range(n):
        p = F.crop(F.pad(Image.fromarray(imgs[i]), 4), int(draw[i, 1]), int(draw[i, 0]), 32, 32)
        if draw[i, 2]:
            p = F.hflip(p)
        p = F.adjust_brightness(p, float(bright[i]))
        ref[i] = F.normalize(F.to_tensor(p), list(dp.CIFAR_MEAN), list(dp.CIFAR_STD)).numpy()
        got = dp.cifar_transform(imgs[i], int(draw[i, 0]), int(draw[i, 1]), bool(draw[i, 2]), float(bright[i]))
        assert np.array_equal(got, ref[i]), ("cifar", i, np.abs(got - ref[i]).max())
    out.update(cifar_img=imgs, cifar_draw=draw, cifar_bright=bright, cifar_out=ref)
    # --- resized crop on 32x32 sources (up-sampling) and on 96x128 sources (down- and up-sampling mixed), + Resize(256)/CenterCrop(224) -----------------
    for tag, (H, W), cnt in (("small", (32, 32), 5), ("large", (300, 400), 3)):
        imgs = rng.integers(0, 256, (cnt, H, W, 3), dtype=np.uint8)
        draw = np.zeros((cnt, 8), dtype=np.int32)
        flip = rng.integers(0, 2, cnt).astype(np.int32)
        res = np.zeros((cnt, 224, 224, 3), dtype=np.uint8)
        for i in range(cnt):
            if i == 0:                     # the test transform: Resize(224) (+ CenterCrop(224))
                d, _ = __import__("libcontinual_b200.data", fromlist=["draw_resize_center"]).draw_resize_center(1, H, W, 224 if tag == "small" else 256, 224)
                draw[i] = d[0]; flip[i] = 0
                p = F.resize(Image.fromarray(imgs[i]), 224 if tag == "small" else 256)
                p = F.center_crop(p, 224)
            else:
                h = int(rng.integers(H // 4, H + 1)); w = int(rng.integers(W // 4, W + 1))
                top = int(rng.integers(0, H - h + 1)); left = int(rng.integers(0, W - w + 1))
                draw[i] = (top, left, h, w, 224, 224, 0, 0)
                p = F.resized_crop(Image.fromarray(imgs[i]), top, left, h, w, (224, 224))
                if flip[i]:
                    p = F.hflip(p)
            res[i] = np.asarray(p)
            t, l, h, w, oh, ow, oy, ox = (int(v) for v in draw[i])
            got = dp.resized_crop_window(imgs[i], t, l, h, w, oh, ow, oy, ox, 224, bool(flip[i]))
            assert np.array_equal(got, F.to_tensor(p).numpy()), (tag, i)
        out.update({f"{tag}_img": imgs, f"{tag}_draw": draw, f"{tag}_flip": flip, f"{tag}_out_u8": res})
    path = os.path.join(ROOT, "tests", "golden", "data_transforms.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
