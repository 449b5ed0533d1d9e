"""TEST INFRASTRUCTURE (not shipped): CPU restatement (numpy) of the reference's per-sample input transforms, parameterised by the random draws so that
a kernel can be compared bit for bit given the same draws.

  cifar_train : core/data/data.py:11-16  RandomCrop(32, padding=4) -> RandomHorizontalFlip -> ColorJitter(brightness=63/255) -> ToTensor -> Normalize
  cifar_test  : core/data/data.py:18     ToTensor -> Normalize
  vit_train   : core/data/data.py:27-31  RandomResizedCrop(224) -> RandomHorizontalFlip -> ToTensor -> Normalize(0, 1)
  vit_test    : core/data/data.py:33-36  Resize(224) -> ToTensor -> Normalize(0, 1)

Pinned against torchvision 0.26 / PIL driven with the same parameters (oracle/make_golden_data.py): both paths are exact — the resize follows PIL's
Resample.c (separable triangle filter, support = max(1, scale), coefficients normalised per output pixel and rounded to 22-bit fixed point, horizontal
pass then vertical pass, uint8 after each)."""
import numpy as np

CIFAR_MEAN = np.array([0.5071, 0.4866, 0.4409], dtype=np.float32)      # data.py:7-8
CIFAR_STD = np.array([0.2675, 0.2565, 0.2761], dtype=np.float32)


def cifar_transform(img, dx=4, dy=4, flip=False, factor=1.0, pad=4, mean=CIFAR_MEAN, std=CIFAR_STD):
    """img uint8 [H, W, 3]; (dx, dy) = left / top of the crop window inside the zero-padded image (4, 4 = no shift); factor = brightness factor
    (PIL ImageEnhance.Brightness = Image.blend(black, img, factor): out = uint8(float32(factor) * v), clipped at 255)."""
    H, W, _ = img.shape
    padded = np.zeros((H + 2 * pad, W + 2 * pad, 3), dtype=np.uint8)
    padded[pad:pad + H, pad:pad + W] = img
    out = padded[dy:dy + H, dx:dx + W]
    if flip:
        out = out[:, ::-1]
    t = np.float32(factor) * out.astype(np.float32)
    out = np.where(t >= 255.0, 255, np.where(t <= 0.0, 0, t.astype(np.int32))).astype(np.uint8) if factor != 1.0 else out
    x = out.astype(np.float32) / np.float32(255.0)
    x = (x - mean) / std
    return np.ascontiguousarray(x.transpose(2, 0, 1))


PRECISION_BITS = 32 - 8 - 2          # PIL Resample.c: 8-bit images are filtered with 22-bit fixed-point coefficients


def _triangle_coeffs(in_size, out_size):
    """PIL `precompute_coeffs` + `normalize_coeffs_8bpc` for the bilinear (triangle) filter over the whole axis: per output index (xmin, int32 coefficients)."""
    scale = in_size / out_size
    fs = max(scale, 1.0)
    support = 1.0 * fs
    out = []
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size)
        ws = np.array([max(0.0, 1.0 - abs((x + xmin - center + 0.5) * (1.0 / fs))) for x in range(xmax - xmin)], dtype=np.float64)
        ww = 0.0
        for v in ws:                       # PIL accumulates the normaliser sequentially
            ww += float(v)
        ws = ws / ww
        out.append((xmin, np.array([int(0.5 + k * (1 << PRECISION_BITS)) for k in ws], dtype=np.int64)))
    return out


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255)


def resized_crop(img, top, left, h, w, out=224, flip=False):
    """img uint8 [H, W, 3] -> float32 [3, out, out] in [0, 1]: F.resized_crop(img, top, left, h, w, (out, out), BILINEAR) -> hflip -> ToTensor, i.e.
    img.crop(box).resize((out, out), BILINEAR): PIL's two passes (horizontal, then vertical), each in 22-bit fixed point with a uint8 result."""
    src = img[top:top + h, left:left + w].astype(np.int64)
    half = 1 << (PRECISION_BITS - 1)
    if w != out:
        tmp = np.zeros((h, out, 3), dtype=np.int64)
        for xx, (x0, ks) in enumerate(_triangle_coeffs(w, out)):
            tmp[:, xx] = _clip8(half + np.tensordot(src[:, x0:x0 + len(ks)], ks, axes=([1], [0])))
    else:
        tmp = src
    if h != out:
        res = np.zeros((out, out, 3), dtype=np.int64)
        for yy, (y0, ks) in enumerate(_triangle_coeffs(h, out)):
            res[yy] = _clip8(half + np.tensordot(tmp[y0:y0 + len(ks)], ks, axes=([0], [0])))
    else:
        res = tmp
    if flip:
        res = res[:, ::-1]
    return np.ascontiguousarray((res.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1))


def resized_crop_window(img, top, left, h, w, oh, ow, oy, ox, out=224, flip=False):
    """Crop box (top, left, h, w) -> resize to (oh, ow) -> the out x out window at (oy, ox) -> flip -> ToTensor: covers RandomResizedCrop (oh = ow = out,
    oy = ox = 0) and Resize + CenterCrop."""
    src = img[top:top + h, left:left + w].astype(np.int64)
    half = 1 << (PRECISION_BITS - 1)
    if w != ow:
        tmp = np.zeros((h, ow, 3), dtype=np.int64)
        for xx, (x0, ks) in enumerate(_triangle_coeffs(w, ow)):
            tmp[:, xx] = _clip8(half + np.tensordot(src[:, x0:x0 + len(ks)], ks, axes=([1], [0])))
    else:
        tmp = src
    if h != oh:
        res = np.zeros((oh, ow, 3), dtype=np.int64)
        for yy, (y0, ks) in enumerate(_triangle_coeffs(h, oh)):
            res[yy] = _clip8(half + np.tensordot(tmp[y0:y0 + len(ks)], ks, axes=([0], [0])))
    else:
        res = tmp
    res = res[oy:oy + out, ox:ox + out]
    if flip:
        res = res[:, ::-1]
    return np.ascontiguousarray((res.astype(np.float32) / np.float32(255.0)).transpose(2, 0, 1))
