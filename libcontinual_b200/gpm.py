"""GPM gradient projection (core/model/gpm.py:65-83, :116-129): after `loss.backward()`, every TRGP layer's weight gradient loses its component inside
the subspace spanned by the stored bases, g <- g - g.view(out, -1) @ (U U^T).  `GPMProjector` keeps the projection matrices of a task as two-term BF16
splits and applies the projection with three tcgen05 GEMMs per layer (`lc_gpm_project_tc`), in place on the gradient tensors.

Only the projection is on the CUDA path so far; the GPM plugin itself (AlexNet_TRGP backbone, basis update by SVD) is not built."""
from __future__ import annotations

from typing import List, Sequence

import torch

from . import _lib
from ._lib import check


class GPMProjector:
    def __init__(self, feature_list: Sequence[torch.Tensor], device="cuda:0"):
        """feature_list[i]: basis U_i [dim_i, r_i] of layer i (gpm.py:124 builds feature_mat[i] = U_i U_i^T from it)."""
        self.lib = _lib.load()
        self.dev = torch.device(device)
        self.hi: List[torch.Tensor] = []
        self.lo: List[torch.Tensor] = []
        self.dims: List[int] = []
        self.err = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._scratch = {}
        st = torch.cuda.current_stream().cuda_stream
        for U in feature_list:
            U = torch.as_tensor(U, dtype=torch.float32, device=self.dev)
            M = (U @ U.T).contiguous()
            M = ((M + M.T) * 0.5).contiguous()                   # exactly symmetric: the kernel reads M as its own transpose
            hi = torch.empty(M.shape, device=self.dev, dtype=torch.bfloat16); lo = torch.empty_like(hi)
            check(self.lib.lc_split_bf16(M.data_ptr(), hi.data_ptr(), lo.data_ptr(), M.numel(), st), "split_bf16")
            self.hi.append(hi); self.lo.append(lo); self.dims.append(M.shape[0])

    def project_(self, i: int, grad: torch.Tensor) -> torch.Tensor:
        """In place: grad <- grad - grad.view(out, -1) @ M_i."""
        assert grad.is_cuda and grad.dtype == torch.float32 and grad.is_contiguous()
        rows, dim = grad.shape[0], grad.numel() // grad.shape[0]
        assert dim == self.dims[i] and dim % 8 == 0
        key = (rows, dim)
        if key not in self._scratch:
            self._scratch[key] = (torch.empty(rows, dim, device=self.dev, dtype=torch.bfloat16), torch.empty(rows, dim, device=self.dev, dtype=torch.bfloat16))
        gh, gl = self._scratch[key]
        check(self.lib.lc_gpm_project_tc(grad.data_ptr(), self.hi[i].data_ptr(), self.lo[i].data_ptr(), rows, dim, gh.data_ptr(), gl.data_ptr(), self.err.data_ptr(),
                                         torch.cuda.current_stream().cuda_stream), "gpm_project_tc")
        return grad
