"""Device-resident input pipeline: the reference's per-sample PIL / torchvision transforms on 24 DataLoader workers followed by an
unpinned synchronous H2D copy of an fp32 batch (core/data/dataset.py:248-266, core/data/data.py:11-35, core/data/dataloader.py:17-38,
config/headers/data.yaml:5-7) become: the uint8 dataset resident in HBM, one small table of random draws per batch crossing PCIe from pinned memory on a
copy stream while the previous step runs, and one launch that gathers + transforms the batch with PIL's / torchvision's own arithmetic
(`lc_augment_cifar_u8`, `lc_resize_crop_u8`).  A 128-image CIFAR batch costs 393 KB of uint8 reads instead of a 1.5 MB fp32 H2D per step.

`GpuLoader` is iterable like the reference's DataLoader (dict batches {'image', 'label'}, `len()`, `.dataset`), so `Trainer._train` / `before_task` /
`after_task` consume it unchanged; images and labels arrive already on the device.

Random draws: torchvision draws from torch's global RNG inside every worker process, a stream nobody can reproduce across loaders.  Here the draws come from
one seeded numpy Generator per loader (restating `RandomCrop.get_params`, `RandomHorizontalFlip`, `ColorJitter.get_params`,
`RandomResizedCrop.get_params` — torchvision/transforms/transforms.py), i.e. the same DISTRIBUTION; given the same draws the pixels are bit-identical to
torchvision's (tests/test_gpu_data.py against tests/golden/data_transforms.npz, written by the real torchvision)."""
from __future__ import annotations

import math
from typing import Dict, Iterator, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import check

CIFAR_MEAN, CIFAR_STD = (0.5071, 0.4866, 0.4409), (0.2675, 0.2565, 0.2761)     # data.py:7-8
VIT_MEAN, VIT_STD = (0.0, 0.0, 0.0), (1.0, 1.0, 1.0)                           # data.py:25-26


# ---- the random draws (host, numpy) --------------------------------------------------------------------------------------------------------------------
def draw_cifar_train(rng: np.random.Generator, B: int, size: int = 32, pad: int = 4, brightness: float = 63 / 255):
    """RandomCrop(size, padding=pad): i, j ~ U{0..2*pad}; RandomHorizontalFlip(p = 0.5); ColorJitter(brightness=b): factor ~ U[max(0, 1-b), 1+b]."""
    draw = np.zeros((B, 4), dtype=np.int32)
    draw[:, 0] = rng.integers(0, 2 * pad + 1, B)          # dx (left)
    draw[:, 1] = rng.integers(0, 2 * pad + 1, B)          # dy (top)
    draw[:, 2] = rng.random(B) < 0.5
    bright = rng.uniform(max(0.0, 1.0 - brightness), 1.0 + brightness, B).astype(np.float32)
    return draw, bright


def draw_identity(B: int, pad: int = 0):
    draw = np.zeros((B, 4), dtype=np.int32)
    draw[:, 0] = pad
    draw[:, 1] = pad
    return draw, None


def draw_resized_crop(rng: np.random.Generator, B: int, H: int, W: int, out: int = 224, scale=(0.08, 1.0), ratio=(3 / 4, 4 / 3)):
    """RandomResizedCrop.get_params: ten attempts at (area ~ U(scale) * H*W, log-ratio ~ U(log ratio)), else the central crop clamped to the ratio bounds;
    RandomHorizontalFlip.  Returns draw [B, 8] = top, left, h, w, OH, OW, oy, ox and flip [B]."""
    draw = np.zeros((B, 8), dtype=np.int32)
    area = H * W
    lr = (math.log(ratio[0]), math.log(ratio[1]))
    for b in range(B):
        box = None
        for _ in range(10):
            ta = area * rng.uniform(scale[0], scale[1])
            ar = math.exp(rng.uniform(lr[0], lr[1]))
            w = int(round(math.sqrt(ta * ar)))
            h = int(round(math.sqrt(ta / ar)))
            if 0 < w <= W and 0 < h <= H:
                box = (int(rng.integers(0, H - h + 1)), int(rng.integers(0, W - w + 1)), h, w)
                break
        if box is None:
            in_ratio = W / H
            if in_ratio < ratio[0]:
                w = W; h = int(round(w / ratio[0]))
            elif in_ratio > ratio[1]:
                h = H; w = int(round(h * ratio[1]))
            else:
                w, h = W, H
            box = ((H - h) // 2, (W - w) // 2, h, w)
        draw[b] = (*box, out, out, 0, 0)
    flip = (rng.random(B) < 0.5).astype(np.int32)
    return draw, flip


def draw_resize_center(B: int, H: int, W: int, resize: int = 224, out: int = 224):
    """Resize(resize) (smaller edge -> resize, torchvision's int() of the longer edge) followed by CenterCrop(out) (identity when the image is already
    out x out): the vit test transform (data.py:33-36) and the ImageNet-R style Resize(256) + CenterCrop(224)."""
    if H <= W:
        oh, ow = resize, int(resize * W / H)
    else:
        oh, ow = int(resize * H / W), resize
    oy, ox = int(round((oh - out) / 2.0)), int(round((ow - out) / 2.0))
    draw = np.tile(np.array([0, 0, H, W, oh, ow, oy, ox], dtype=np.int32), (B, 1))
    return draw, None


# ---- dataset / loader ------------------------------------------------------------------------------------------------------------------------------------
class DeviceImageDataset:
    """uint8 images [N, H, W, 3] and int64 labels resident on the device.  `.images` / `.labels` keep host views for the hooks that read
    `train_loader.dataset` (herding, linearherdingbuffer.py:94-111)."""

    def __init__(self, images_u8, labels, device="cuda:0"):
        images_u8 = torch.as_tensor(images_u8)
        assert images_u8.dtype == torch.uint8 and images_u8.dim() == 4 and images_u8.shape[-1] == 3, "images: uint8 [N, H, W, 3]"
        self.device = torch.device(device)
        self.images = images_u8.cpu().numpy()
        self.labels = np.asarray(torch.as_tensor(labels).cpu().numpy(), dtype=np.int64)
        self.images_dev = images_u8.to(self.device).contiguous()
        self.labels_dev = torch.as_tensor(self.labels).to(self.device)
        self.N, self.H, self.W = int(images_u8.shape[0]), int(images_u8.shape[1]), int(images_u8.shape[2])

    def __len__(self):
        return self.N


class GpuLoader:
    """Iterable over dict batches {'image': fp32 [B, 3, h, w] on the device, 'label': int64 [B] on the device}.

    transform: 'cifar_train' | 'cifar_test' | 'vit_train' | 'vit_test' (core/data/data.py:11-36).  The draws of batch i+1 are generated on the host and
    copied from pinned memory on a copy stream while batch i is being consumed (double-buffered)."""

    def __init__(self, dataset: DeviceImageDataset, batch_size: int, transform: str = "cifar_train", shuffle: bool = True, drop_last: bool = False,
                 seed: int = 0, out_size: int = 224):
        if transform not in ("cifar_train", "cifar_test", "vit_train", "vit_test"):
            raise ValueError(f"unknown transform {transform!r}")
        self.dataset, self.batch_size, self.transform, self.shuffle, self.drop_last = dataset, int(batch_size), transform, shuffle, drop_last
        self.rng = np.random.default_rng(seed)
        self.out_size = out_size
        self.lib = _lib.load()
        dev = dataset.device
        self.dev = dev
        self.copy = torch.cuda.Stream(device=dev)
        vit = transform.startswith("vit")
        mean, std = (VIT_MEAN, VIT_STD) if vit else (CIFAR_MEAN, CIFAR_STD)
        self.mean = np.asarray(mean, dtype=np.float32)
        self.std = np.asarray(std, dtype=np.float32)
        B = self.batch_size
        nd = 8 if vit else 4
        # double-buffered pinned staging of (indices, draws, flip / brightness) and their device copies
        self.h_idx = [torch.empty(B, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.h_draw = [torch.empty(B, nd, dtype=torch.int32).pin_memory() for _ in range(2)]
        self.h_aux = [torch.empty(B, dtype=torch.int32 if vit else torch.float32).pin_memory() for _ in range(2)]
        self.d_idx = [torch.empty(B, dtype=torch.int64, device=dev) for _ in range(2)]
        self.d_draw = [torch.empty(B, nd, dtype=torch.int32, device=dev) for _ in range(2)]
        self.d_aux = [torch.empty(B, dtype=torch.int32 if vit else torch.float32, device=dev) for _ in range(2)]
        self.ev = [torch.cuda.Event() for _ in range(2)]
        self.ev_used = [torch.cuda.Event() for _ in range(2)]      # consumer finished reading the device tables of slot k
        self.used_rec = [False, False]
        if vit:
            nbytes = int(self.lib.lc_resize_scratch_bytes(B, dataset.H, out_size))
            self.scratch = torch.empty((nbytes + 15) // 16 * 16, dtype=torch.uint8, device=dev)
        self._last_draw = None

    def __len__(self):
        n = len(self.dataset)
        return n // self.batch_size if self.drop_last else (n + self.batch_size - 1) // self.batch_size

    def _draws(self, B):
        ds = self.dataset
        if self.transform == "cifar_train":
            return draw_cifar_train(self.rng, B, ds.H, 4)
        if self.transform == "cifar_test":
            return draw_identity(B, 4)
        if self.transform == "vit_train":
            return draw_resized_crop(self.rng, B, ds.H, ds.W, self.out_size)
        return draw_resize_center(B, ds.H, ds.W, self.out_size, self.out_size)

    def _stage(self, slot: int, idx: np.ndarray):
        B = len(idx)
        draw, aux = self._draws(B)
        if self.used_rec[slot]:
            self.ev[slot].synchronize()          # the previous copy out of this pinned slot has left the host before it is overwritten
        self.h_idx[slot][:B].copy_(torch.from_numpy(idx))
        self.h_draw[slot][:B].copy_(torch.from_numpy(draw))
        has_aux = aux is not None
        if has_aux:
            self.h_aux[slot][:B].copy_(torch.from_numpy(aux))
        with torch.cuda.stream(self.copy):
            if self.used_rec[slot]:
                self.copy.wait_event(self.ev_used[slot])
            self.d_idx[slot][:B].copy_(self.h_idx[slot][:B], non_blocking=True)
            self.d_draw[slot][:B].copy_(self.h_draw[slot][:B], non_blocking=True)
            if has_aux:
                self.d_aux[slot][:B].copy_(self.h_aux[slot][:B], non_blocking=True)
            self.ev[slot].record(self.copy)
        return B, has_aux, (draw, aux)

    def _launch(self, slot: int, B: int, has_aux: bool) -> Dict[str, torch.Tensor]:
        ds = self.dataset
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_event(self.ev[slot])
        st = cur.cuda_stream
        aux = self.d_aux[slot].data_ptr() if has_aux else None
        if self.transform.startswith("cifar"):
            out = torch.empty(B, 3, ds.H, ds.W, device=self.dev)
            check(self.lib.lc_augment_cifar_u8(ds.images_dev.data_ptr(), self.d_idx[slot].data_ptr(), self.d_draw[slot].data_ptr(), aux, out.data_ptr(), B, ds.H,
                                               ds.W, 4, self.mean.ctypes.data, self.std.ctypes.data, st), "lc_augment_cifar_u8")
        else:
            out = torch.empty(B, 3, self.out_size, self.out_size, device=self.dev)
            check(self.lib.lc_resize_crop_u8(ds.images_dev.data_ptr(), self.d_idx[slot].data_ptr(), self.d_draw[slot].data_ptr(), aux, out.data_ptr(), B, ds.H,
                                             ds.W, self.out_size, self.mean.ctypes.data, self.std.ctypes.data, self.scratch.data_ptr(), st), "lc_resize_crop_u8")
        label = ds.labels_dev.index_select(0, self.d_idx[slot][:B])
        self.ev_used[slot].record(cur)
        self.used_rec[slot] = True
        return {"image": out, "label": label}

    def __iter__(self) -> Iterator[Dict[str, torch.Tensor]]:
        n = len(self.dataset)
        order = self.rng.permutation(n) if self.shuffle else np.arange(n)
        nb = len(self)
        if nb == 0:
            return
        chunks = [order[i * self.batch_size:min(n, (i + 1) * self.batch_size)].astype(np.int64) for i in range(nb)]
        pending = self._stage(0, chunks[0])
        for i in range(nb):
            slot = i & 1
            B, has_aux, draws = pending
            if i + 1 < nb:
                pending = self._stage(slot ^ 1, chunks[i + 1])       # the next batch's tables cross PCIe while this batch is consumed
            self._last_draw = (chunks[i], *draws)
            yield self._launch(slot, B, has_aux)


def transform_batch(images_u8: torch.Tensor, transform: str, draw: np.ndarray, aux: Optional[np.ndarray] = None, out_size: int = 224) -> torch.Tensor:
    """One batch of device uint8 images [B, H, W, 3] through a transform with explicit draws (the entry the parity tests use)."""
    lib = _lib.load()
    dev = images_u8.device
    B, H, W, _ = images_u8.shape
    vit = transform.startswith("vit")
    mean = np.asarray(VIT_MEAN if vit else CIFAR_MEAN, dtype=np.float32)
    std = np.asarray(VIT_STD if vit else CIFAR_STD, dtype=np.float32)
    d_draw = torch.from_numpy(np.ascontiguousarray(draw, dtype=np.int32)).to(dev)
    st = torch.cuda.current_stream(dev).cuda_stream
    if not vit:
        d_aux = torch.from_numpy(np.ascontiguousarray(aux, dtype=np.float32)).to(dev) if aux is not None else None
        out = torch.empty(B, 3, H, W, device=dev)
        check(lib.lc_augment_cifar_u8(images_u8.data_ptr(), None, d_draw.data_ptr(), d_aux.data_ptr() if d_aux is not None else None, out.data_ptr(), B, H, W, 4,
                                      mean.ctypes.data, std.ctypes.data, st), "lc_augment_cifar_u8")
        return out
    d_aux = torch.from_numpy(np.ascontiguousarray(aux, dtype=np.int32)).to(dev) if aux is not None else None
    scratch = torch.empty((int(lib.lc_resize_scratch_bytes(B, H, out_size)) + 15) // 16 * 16, dtype=torch.uint8, device=dev)
    out = torch.empty(B, 3, out_size, out_size, device=dev)
    check(lib.lc_resize_crop_u8(images_u8.data_ptr(), None, d_draw.data_ptr(), d_aux.data_ptr() if d_aux is not None else None, out.data_ptr(), B, H, W, out_size,
                                mean.ctypes.data, std.ctypes.data, scratch.data_ptr(), st), "lc_resize_crop_u8")
    return out
