"""Exemplar memory for iCaRL on device tensors — mirror of `LinearHerdingBuffer` (core/model/buffer/linearherdingbuffer.py:10-163)
for tensor-backed task data (the reference keeps image PATHS and re-decodes them).  Selection runs in
`lc_herding_select`, so the chosen indices are bit-identical to the reference's greedy loop on the same features."""
from __future__ import annotations

from typing import List

import torch

from ._lib import check, load


class HerdingBuffer:
    def __init__(self, buffer_size: int, batch_size: int = 32, **kwargs):
        self.buffer_size, self.batch_size = buffer_size, batch_size
        self.images: List[torch.Tensor] = []      # per-class exemplar tensors, class-ordered
        self.labels: List[torch.Tensor] = []
        self.total_classes = 0

    def is_empty(self) -> bool:
        return len(self.labels) == 0

    def get_all_data(self):
        return torch.cat(self.images), torch.cat(self.labels)

    def reduce_old_data(self, task_idx: int, total_cls_num: int) -> None:
        """Keep the first `buffer_size // total_cls_num` (herding-ordered) exemplars of every stored class (:55-75)."""
        per = max(1, self.buffer_size // total_cls_num)
        if task_idx > 0:
            self.images = [x[:per] for x in self.images]
            self.labels = [y[:per] for y in self.labels]

    @torch.no_grad()
    def _features(self, model, x: torch.Tensor, transform=None) -> torch.Tensor:
        """L2-normalised eval-mode features of raw items `x`, each batch passed through `transform` (the validation transform, :97) first."""
        bb = model.backbone
        was = bb.training
        bb.eval()
        out = []
        for i in range(0, x.shape[0], 32):                 # batch 32, no shuffle (:118-126)
            xb = x[i:i + 32]
            f = bb(xb if transform is None else transform(xb))["features"]
            out.append(f / f.norm(dim=1).view(-1, 1))
        bb.train(was)
        return torch.cat(out)

    @torch.no_grad()
    def herding_indices(self, model, x: torch.Tensor, y: torch.Tensor, per_class: int, transform=None) -> torch.Tensor:
        """x/y: the current task's samples, sorted by class.  Returns the selected global indices (class-major)."""
        lib = load()
        dev = model.engine.device
        feats = self._features(model, x.to(dev), transform).contiguous()
        classes, counts = torch.unique_consecutive(y.cpu(), return_counts=True)
        assert torch.equal(classes, torch.sort(classes)[0]) and len(torch.unique(classes)) == len(classes), "task data must be sorted by class"
        begins = torch.zeros(len(classes) + 1, dtype=torch.int32)
        begins[1:] = torch.cumsum(counts, 0)
        begins = begins.to(dev)
        out = torch.empty(len(classes), per_class, dtype=torch.int64, device=dev)
        work = torch.empty_like(feats)
        check(lib.lc_herding_select(feats.data_ptr(), begins.data_ptr(), len(classes), feats.shape[1], per_class, work.data_ptr(), out.data_ptr(),
                                    torch.cuda.current_stream().cuda_stream), "lc_herding_select")
        return out

    def update(self, model, x: torch.Tensor, y: torch.Tensor, total_cls_num: int, transform=None) -> torch.Tensor:
        """x: RAW task items (stored as they are, like the reference stores paths); `transform` only shapes what the feature pass sees."""
        per = max(1, self.buffer_size // total_cls_num)
        idx = self.herding_indices(model, x, y, per, transform).cpu()
        for row in idx:
            row = row[row >= 0]
            self.images.append(x[row])
            self.labels.append(y[row])
        self.total_classes = total_cls_num
        return idx

    @torch.no_grad()
    def class_means(self, model, transform=None) -> torch.Tensor:
        """`ICarl.calc_class_mean` (core/model/icarl.py:226-287): per class mean of the L2-normalised exemplar features, re-normalised."""
        dev = model.engine.device
        means = []
        for x in self.images:
            f = self._features(model, x.to(dev), transform)
            m = f.mean(0)
            means.append(m / m.norm())
        return torch.stack(means)
