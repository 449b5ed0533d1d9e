"""Device-side state and kernel sequencing for the backbones built on the generic layer kernels (`csrc/nn_ops.cuh`) and the tcgen05 GEMM:

* `ResNet18Engine` — torchvision-style ResNet18 of the reference (`core/model/backbone/resnet.py:26-64,110-246`, factory `resnet18` :259-267) with the
  tiny-imagenet / cifar stems (:133-150), for LwF at 64 x 64 (`lwf.py:52-70`, BASELINE config C5).  It exposes the same attributes and methods as
  `engine.ResNetEngine`, so `model.LWF / EWC / ICarl / Finetune`, `optim.SGD` and `trainer.GraphedStep` drive it unchanged.
* `AlexNetEngine` — `AlexNet_TRGP` (`core/model/backbone/alexnet.py:94-156`) for GPM (`gpm.py:45-204`).

Arithmetic: every convolution / linear layer is a BF16-operand, fp32-accumulate tcgen05 GEMM — 3x3 / 1x1 convolutions with Cin % 64 == 0 as IMPLICIT GEMMs
(the producer warp reads the NHWC activation through a rank-4 tensor map, tap = coordinate offset, padding = out-of-bounds zero fill), the rest through an
explicit BF16 patch matrix.  BatchNorm statistics, the residual stream, losses, gradients of the parameters and the optimizer are fp32.
Everything here is launch plumbing: torch supplies device memory and streams, every FLOP runs in `liblc_b200.so`.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, GemmDesc, check, stream_ptr

KORDER_TAP_C, KORDER_C_TAP = 0, 1
SRC_NHWC_BF16, SRC_NCHW_F32, SRC_NHWC_F32 = 0, 1, 2
BN_EPS, BN_MOMENTUM = 1e-5, 0.1
FLAT_SCRATCH_FLOATS = 2 * 296 + 8


def _up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _p(t, off_elems: int = 0):
    """Device pointer of a tensor (+ element offset), or NULL."""
    if t is None:
        return None
    return t.data_ptr() + off_elems * t.element_size()


class NNOps:
    """Thin ctypes wrappers (one launch sequence each) + launch counting."""

    def __init__(self, device=None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.LcError("libcontinual_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            check(self.lib.lc_device_check(), "lc_device_check: built for sm_100a only")
        self.err = torch.zeros(4, dtype=torch.int32, device=self.device)
        self.bn_scratch = torch.zeros(int(self.lib.lc_nn_bn_scratch_floats(2048)), device=self.device)
        self.launches = 0

    def tensor_core_error(self) -> bool:
        return bool(int(self.err[0]) != 0)

    # ---- GEMMs ---------------------------------------------------------------------------------------------------------------------------
    def gemm(self, A, lda, B, ldb, C, ldc, M, N, K, *, out_f32=True, bias=None, residual=None, ldr=0, ksplit=0, stride_split=0):
        """C[M][N] = A[M][K] B[N][K]^T (BF16 operands given as device pointers / tensors)."""
        d = GemmDesc(A=A, lda=lda, B=B, ldb=ldb, C=C, ldc=ldc, bias=bias, residual=residual, ldr=ldr, M=M, N=N, K=K, batch_in=1, batch_out=1,
                     out_f32=int(out_f32), alpha=1.0, ksplit=ksplit, strideC_split=stride_split)
        check(self.lib.lc_gemm_bf16_ex(ctypes.byref(d), self.err.data_ptr(), stream_ptr()), f"lc_gemm_bf16_ex M={M} N={N} K={K}")
        self.launches += 1

    def conv(self, X, Wk, Y, N, H, W, C, Cout, ks, stride, pad, *, out_f32=True, residual=None):
        Ho, Wo = (H + 2 * pad - ks) // stride + 1, (W + 2 * pad - ks) // stride + 1
        d = ConvDesc(X=X, Wk=Wk, Y=Y, bias=None, residual=residual, ldc=Cout, ldr=Cout, N=N, H=H, W=W, C=C, Cout=Cout, ks=ks, stride=stride, pad=pad,
                     Ho=Ho, Wo=Wo, out_f32=int(out_f32))
        check(self.lib.lc_conv_gemm_bf16(ctypes.byref(d), self.err.data_ptr(), stream_ptr()), f"lc_conv_gemm_bf16 {N}x{H}x{W}x{C}->{Cout} k{ks}s{stride}")
        self.launches += 1

    def wgrad_splits(self, cout: int, kcols: int, kdim: int) -> int:
        """Split-K factor of a weight-gradient GEMM [cout][kcols] contracting over kdim: enough CTAs for the machine, >= 4 K blocks each."""
        tiles = ((cout + 127) // 128) * ((kcols + 255) // 256 if kcols > 128 else 1)
        nkb = (kdim + 63) // 64
        s = max(1, min((148 + tiles - 1) // tiles, nkb // 4, 64))
        kper = (nkb + s - 1) // s
        return (nkb + kper - 1) // kper

    # ---- layer kernels ---------------------------------------------------------------------------------------------------------------------
    def im2col(self, src, kind, N, H, W, C, ks, stride, pad, korder, col, ld_col, colT, ld_colT, Kp):
        check(self.lib.lc_nn_im2col(src, kind, N, H, W, C, ks, stride, pad, korder, col, ld_col, colT, ld_colT, Kp, stream_ptr()), "lc_nn_im2col")
        self.launches += 1

    def col2im(self, dcol, ld, addend, dx, N, H, W, C, ks, stride, pad, korder):
        check(self.lib.lc_nn_col2im(dcol, ld, addend, dx, N, H, W, C, ks, stride, pad, korder, stream_ptr()), "lc_nn_col2im")
        self.launches += 1

    def bn_stats(self, y, M, C, gamma, beta, running, aff):
        check(self.lib.lc_nn_bn_stats(y, M, C, gamma, beta, BN_EPS, BN_MOMENTUM, running, aff, self.bn_scratch.data_ptr(), stream_ptr()), "lc_nn_bn_stats")
        self.launches += 2

    def bn_eval(self, running, C, gamma, beta, aff):
        check(self.lib.lc_nn_bn_eval_affine(running, C, gamma, beta, BN_EPS, aff, stream_ptr()), "lc_nn_bn_eval_affine")
        self.launches += 1

    def bn_act(self, y, aff, M, C, *, res=None, res_aff=None, relu=True, drop_p=0.0, rng=None, rng_stream=0, out_bf16=None, out_f32=None):
        check(self.lib.lc_nn_bn_act(y, aff, res, res_aff, M, C, int(relu), float(drop_p), rng, rng_stream, out_bf16, out_f32, stream_ptr()), "lc_nn_bn_act")
        self.launches += 1

    def bn_bwd(self, g, y, aff, M, C, *, act_f32=None, act_bf16=None, gscale=1.0, dgamma=None, dbeta=None, dy_bf16=None, dy_f32=None, dz_out=None):
        check(self.lib.lc_nn_bn_backward(g, act_f32, act_bf16, float(gscale), y, aff, M, C, dgamma, dbeta, dy_bf16, dy_f32, dz_out, self.bn_scratch.data_ptr(),
                                         stream_ptr()), "lc_nn_bn_backward")
        self.launches += 3

    def pack(self, w, cout, cin, ks, korder, mode, out, ld):
        check(self.lib.lc_nn_pack_weight(w, cout, cin, ks, korder, mode, out, ld, stream_ptr()), "lc_nn_pack_weight")
        self.launches += 1

    def wgrad_reduce(self, partial, nsplit, cout, cin, ks, korder, ldp, dw):
        check(self.lib.lc_nn_wgrad_reduce(partial, nsplit, cout, cin, ks, korder, ldp, dw, stream_ptr()), "lc_nn_wgrad_reduce")
        self.launches += 1

    def transpose_bf16(self, src, ld_in, rows, cols, dst, ld_out):
        check(self.lib.lc_transpose_bf16(src, ld_in, rows, cols, dst, ld_out, stream_ptr()), "lc_transpose_bf16")
        self.launches += 1

    def cast_transpose(self, src, rows, cols, out, ld, outT, ldT):
        check(self.lib.lc_nn_cast_transpose(src, rows, cols, out, ld, outT, ldT, stream_ptr()), "lc_nn_cast_transpose")
        self.launches += 1


# ======================================================================================================================================================
# ResNet18
# ======================================================================================================================================================
def resnet18_param_layout(in_ch: int = 3):
    """(name, shape) of every backbone parameter in the reference's registration order (resnet.py:133-156 stem `conv1` Sequential, `_make_layer`
    :196-218, BasicBlock :26-46), BN prefixes in buffer order, and the conv specs (name, bn, cin, cout, ks, stride, pad)."""
    params = [("conv1.0.weight", (64, in_ch, 3, 3)), ("conv1.1.weight", (64,)), ("conv1.1.bias", (64,))]
    bns = ["conv1.1"]
    convs = [("conv1.0", "conv1.1", in_ch, 64, 3, 1, 1)]
    inpl = 64
    for li, (planes, stride) in enumerate([(64, 1), (128, 2), (256, 2), (512, 2)], start=1):
        for b in range(2):
            pre = f"layer{li}.{b}"
            s, cin = (stride, inpl) if b == 0 else (1, planes)
            params += [(pre + ".conv1.weight", (planes, cin, 3, 3)), (pre + ".bn1.weight", (planes,)), (pre + ".bn1.bias", (planes,)),
                       (pre + ".conv2.weight", (planes, planes, 3, 3)), (pre + ".bn2.weight", (planes,)), (pre + ".bn2.bias", (planes,))]
            bns += [pre + ".bn1", pre + ".bn2"]
            convs += [(pre + ".conv1", pre + ".bn1", cin, planes, 3, s, 1), (pre + ".conv2", pre + ".bn2", planes, planes, 3, 1, 1)]
            if b == 0 and (s != 1 or cin != planes):
                params += [(pre + ".downsample.0.weight", (planes, cin, 1, 1)), (pre + ".downsample.1.weight", (planes,)), (pre + ".downsample.1.bias", (planes,))]
                bns.append(pre + ".downsample.1")
                convs.append((pre + ".downsample.0", pre + ".downsample.1", cin, planes, 1, s, 0))
        inpl = planes
    return params, bns, convs


class _Conv:
    """One convolution + its BatchNorm: geometry, arena offsets, packed weights."""
    __slots__ = ("name", "bn", "cin", "cout", "ks", "stride", "pad", "H", "Ho", "w_off", "g_off", "b_off", "r_off", "K", "Kp", "wk", "wd", "wd_mode", "implicit",
                 "nsplit", "ldp")


class _R18Workspace:
    """Activations of one replica (student: everything the backward needs; teacher: the same buffers, simply not read again)."""

    def __init__(self, eng: "ResNet18Engine", B: int):
        dev, f32, bf = eng.device, torch.float32, torch.bfloat16
        e = lambda *s, dt=f32: torch.empty(*s, device=dev, dtype=dt)
        self.B = B
        self.y: Dict[str, torch.Tensor] = {}
        self.aff: Dict[str, torch.Tensor] = {}
        for c in eng.convs:
            self.y[c.name] = e(B * c.Ho * c.Ho, c.cout)
            self.aff[c.name] = torch.zeros(4 * c.cout, device=dev)
        s = eng.convs[0]
        M0 = B * s.Ho * s.Ho
        self.col0 = e(M0, s.Kp, dt=bf)
        self.colT0 = e(s.Kp, _up(M0, 8), dt=bf)
        self.a0 = e(M0, 64)
        if eng.maxpool:
            self.pool_idx = torch.empty(B * eng.H1 * eng.H1, 64, device=dev, dtype=torch.uint8)
        self.x0_f32, self.x0_bf = e(B * eng.H1 * eng.H1, 64), e(B * eng.H1 * eng.H1, 64, dt=bf)
        self.a1: Dict[str, torch.Tensor] = {}
        self.out_f32: Dict[str, torch.Tensor] = {}
        self.out_bf: Dict[str, torch.Tensor] = {}
        for blk in eng.blocks:
            c2 = blk["conv2"]
            M = B * c2.Ho * c2.Ho
            self.a1[blk["name"]] = e(M, c2.cout, dt=bf)
            self.out_f32[blk["name"]] = e(M, c2.cout)
            self.out_bf[blk["name"]] = e(M, c2.cout, dt=bf)
        self.feat = e(B, 512)
        # BF16 forward operands of every filter.  The student's are rebuilt every forward (its fp32 master weights move every step); a frozen teacher's
        # are packed once (`packed_for` remembers which parameter arena they came from) — and the two never share a buffer, since the teacher's forward
        # runs concurrently on a second stream
        self.wk = {c.name: torch.zeros(c.cout, c.Kp, device=dev, dtype=bf) for c in eng.convs}
        self.packed_for = None


class ResNet18Engine(NNOps):
    def __init__(self, max_batch: int = 256, num_class_cap: int = 200, device=None, in_ch: int = 3, img: int = 64, maxpool: bool = True):
        super().__init__(device)
        assert in_ch == 3 and img in (32, 64), "stems of resnet.py:133-150 at 32x32 / 64x64"
        dev = self.device
        self.max_batch, self.cap, self.feat_dim, self.img, self.in_ch, self.maxpool = max_batch, num_class_cap, 512, img, in_ch, maxpool
        self.H1 = img // 2 if maxpool else img
        assert self.H1 * self.H1 % 128 == 0, "layer1 resolution must tile into 128-pixel GEMM rows"
        self.layout, self.bn_names, specs = resnet18_param_layout(in_ch)
        self.param_off: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape in self.layout:
            n = 1
            for d in shape:
                n *= d
            self.param_off[name] = (off, shape)
            off += n
        self.n_backbone = off
        self.off_fc_w = _up(off, 4)
        self.off_fc_b = self.off_fc_w + self.cap * self.feat_dim
        self.n_total = _up(self.off_fc_b + self.cap, 4)
        self.params = torch.zeros(self.n_total, device=dev)
        self.grads = torch.zeros(self.n_total, device=dev)
        self.rstat_off: Dict[str, Tuple[int, int]] = {}
        ro = 0
        for bn in self.bn_names:
            c = self.param_off[bn + ".weight"][1][0]
            self.rstat_off[bn] = (ro, c)
            ro += 2 * c
        self.rstat = torch.zeros(ro, device=dev)
        self.reset_running_stats()
        # conv plan
        self.convs: List[_Conv] = []
        H = img
        by_name: Dict[str, _Conv] = {}
        for (name, bn, cin, cout, ks, stride, pad) in specs:
            c = _Conv()
            c.name, c.bn, c.cin, c.cout, c.ks, c.stride, c.pad = name, bn, cin, cout, ks, stride, pad
            if name == "conv1.0":
                c.H = img
            elif name.endswith("conv1"):
                c.H = H
            elif name.endswith("downsample.0"):
                c.H = by_name[name[:-len("downsample.0")] + "conv1"].H       # the shortcut reads the block INPUT
            else:
                c.H = by_name[name[:-1] + "1"].Ho
            c.Ho = (c.H + 2 * pad - ks) // stride + 1
            c.w_off, c.g_off, c.b_off = self.param_off[name + ".weight"][0], self.param_off[bn + ".weight"][0], self.param_off[bn + ".bias"][0]
            c.r_off = self.rstat_off[bn][0]
            c.K = cin * ks * ks
            c.Kp = _up(c.K, 8)
            c.implicit = cin % 64 == 0
            c.wk = None
            if c.implicit and stride == 1 and ks == 3:
                c.wd_mode, c.wd = 2, torch.zeros(cin, 9 * cout, device=dev, dtype=torch.bfloat16)       # flipped: data gradient as a convolution of dY
            elif name != "conv1.0":
                c.wd_mode, c.wd = 1, torch.zeros(c.K, cout, device=dev, dtype=torch.bfloat16)           # transposed: dcol = dY * W, then col2im
            else:
                c.wd_mode, c.wd = -1, None
            self.convs.append(c)
            by_name[name] = c
            if name == "conv1.0":
                H = self.H1
            elif name.endswith("conv2"):
                H = c.Ho
        self.conv_by_name = by_name
        self.blocks = []
        for li in range(1, 5):
            for b in range(2):
                pre = f"layer{li}.{b}"
                self.blocks.append({"name": pre, "conv1": by_name[pre + ".conv1"], "conv2": by_name[pre + ".conv2"], "down": by_name.get(pre + ".downsample.0")})
        B = max_batch
        for c in self.convs:
            M = B * c.Ho * c.Ho
            c.nsplit = self.wgrad_splits(c.cout, c.Kp, M)
            c.ldp = _up(c.Kp, 4)
        self.ws = _R18Workspace(self, B)
        # backward scratch (sized for the largest layer)
        f32, bf = torch.float32, torch.bfloat16
        e = lambda n, dt=f32: torch.empty(n, device=dev, dtype=dt)
        gmax = max(B * c.Ho * c.Ho * c.cout for c in self.convs[1:])
        self.G = [e(gmax), e(gmax)]
        self.Gz, self.dA, self.dXd = e(gmax), e(gmax), e(gmax)
        self.dy_bf = e(max(B * c.Ho * c.Ho * c.cout for c in self.convs), bf)
        self.dyT = e(max(c.cout * _up(B * c.Ho * c.Ho, 8) for c in self.convs), bf)
        self.colT = e(max(c.Kp * _up(B * c.Ho * c.Ho, 8) for c in self.convs), bf)
        self.dcol = e(max([B * c.Ho * c.Ho * c.K for c in self.convs if c.wd_mode == 1] + [8]), bf)
        self.wpart = e(max(c.nsplit * c.cout * c.ldp for c in self.convs))
        self.dA0 = e(B * img * img * 64)
        # head / loss state (same names as ResNetEngine: the method classes and the flat optimizer are shared)
        self.logits = torch.zeros(B, self.cap, device=dev)
        self.dlogits = torch.zeros(B, self.cap, device=dev)
        self.pred = torch.zeros(B, dtype=torch.int64, device=dev)
        self.scal = torch.zeros(8, device=dev)
        self.dfeat = torch.zeros(B, self.feat_dim, device=dev)
        self.flat_scratch = torch.zeros(FLAT_SCRATCH_FLOATS, device=dev)
        self.flat_counter = torch.zeros(4, dtype=torch.int32, device=dev)
        self.hp = torch.zeros(16, device=dev)
        self._lamda_dev = None
        self.autograd_grads = None
        self.ncls = 0
        self.precision = "bf16"
        self._packed_version = None

    # ---- views (ResNetEngine interface) ------------------------------------------------------------------------------------------------------
    def param_view(self, name: str, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, shape = self.param_off[name]
        n = 1
        for d in shape:
            n *= d
        return (self.params if arena is None else arena)[off:off + n].view(shape)

    def fc_views(self, ncls: int, arena: Optional[torch.Tensor] = None):
        a = self.params if arena is None else arena
        return (a[self.off_fc_w:self.off_fc_w + ncls * self.feat_dim].view(ncls, self.feat_dim), a[self.off_fc_b:self.off_fc_b + ncls])

    def running_views(self, bn: str, rstat: Optional[torch.Tensor] = None):
        off, c = self.rstat_off[bn]
        r = self.rstat if rstat is None else rstat
        return r[off:off + c], r[off + c:off + 2 * c]

    def reset_running_stats(self):
        self.rstat.zero_()
        for bn in self.bn_names:
            self.running_views(bn)[1].fill_(1.0)

    def set_precision(self, precision: str):
        if precision not in ("bf16", "tc"):
            raise ValueError("ResNet18Engine computes its contractions with BF16 operands on tcgen05 (fp32 accumulation): precision must be 'bf16' (alias 'tc')")

    def features(self, batch: int, ws=None) -> torch.Tensor:
        return (self.ws if ws is None else ws).feat[:batch]

    def fmaps(self, batch: int):
        outs = []
        for li in range(1, 5):
            blk = self.blocks[2 * li - 1]
            c = blk["conv2"]
            outs.append(self.ws.out_f32[blk["name"]][:batch * c.Ho * c.Ho].view(batch, c.Ho, c.Ho, c.cout).permute(0, 3, 1, 2))
        return outs

    # ---- forward -------------------------------------------------------------------------------------------------------------------------------
    def _pack_weights(self, params: torch.Tensor, ws: "_R18Workspace", frozen: bool):
        """BF16 GEMM operands of every filter: every forward for live weights, once for a frozen teacher."""
        if frozen and ws.packed_for == params.data_ptr():
            return
        for c in self.convs:
            self.pack(_p(params, c.w_off), c.cout, c.cin, c.ks, KORDER_TAP_C, 0, ws.wk[c.name].data_ptr(), c.Kp)
        ws.packed_for = params.data_ptr() if frozen else None

    def _pack_weights_bwd(self, params: torch.Tensor):
        for c in self.convs:
            if c.wd_mode == 2:
                self.pack(_p(params, c.w_off), c.cout, c.cin, c.ks, KORDER_TAP_C, 2, c.wd.data_ptr(), 9 * c.cout)
            elif c.wd_mode == 1:
                self.pack(_p(params, c.w_off), c.cout, c.cin, c.ks, KORDER_TAP_C, 1, c.wd.data_ptr(), c.cout)

    def _conv_fwd(self, c: _Conv, x_bf, B: int, ws: _R18Workspace):
        self.conv(x_bf.data_ptr(), ws.wk[c.name].data_ptr(), ws.y[c.name].data_ptr(), B, c.H, c.H, c.cin, c.cout, c.ks, c.stride, c.pad)

    def _bn(self, c: _Conv, B: int, params, rstat, ws: _R18Workspace, train: bool, update_running: bool):
        M = B * c.Ho * c.Ho
        run = _p(rstat, c.r_off) if rstat is not None else None
        if train:
            self.bn_stats(ws.y[c.name].data_ptr(), M, c.cout, _p(params, c.g_off), _p(params, c.b_off), run if update_running else None, ws.aff[c.name].data_ptr())
        else:
            self.bn_eval(run, c.cout, _p(params, c.g_off), _p(params, c.b_off), ws.aff[c.name].data_ptr())

    def forward(self, x: torch.Tensor, train: bool, update_running: bool = True, params=None, rstat=None, ws=None):
        B = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape[1:]) == (self.in_ch, self.img, self.img) and B <= self.max_batch
        params = self.params if params is None else params
        rstat = self.rstat if rstat is None else rstat
        ws = self.ws if ws is None else ws
        self._pack_weights(params, ws, frozen=ws is not self.ws)
        # stem (resnet.py:145-150): 3x3 conv on 3 channels through an explicit patch matrix (K = 27), BN, ReLU, MaxPool(3, 2, 1)
        s = self.convs[0]
        M0 = B * s.Ho * s.Ho
        self.im2col(x.data_ptr(), SRC_NCHW_F32, B, self.img, self.img, 3, 3, 1, 1, KORDER_TAP_C, ws.col0.data_ptr(), s.Kp, ws.colT0.data_ptr() if train else None,
                    _up(M0, 8), s.Kp)
        self.gemm(ws.col0.data_ptr(), s.Kp, ws.wk[s.name].data_ptr(), s.Kp, ws.y[s.name].data_ptr(), 64, M0, 64, s.Kp)
        self._bn(s, B, params, rstat, ws, train, update_running)
        if self.maxpool:
            self.bn_act(ws.y[s.name].data_ptr(), ws.aff[s.name].data_ptr(), M0, 64, out_f32=ws.a0.data_ptr())
            check(self.lib.lc_nn_maxpool_forward(ws.a0.data_ptr(), B, self.img, self.img, 64, 3, 2, 1, ws.x0_f32.data_ptr(), ws.x0_bf.data_ptr(), ws.pool_idx.data_ptr(),
                                                 stream_ptr()), "lc_nn_maxpool_forward")
            self.launches += 1
        else:
            self.bn_act(ws.y[s.name].data_ptr(), ws.aff[s.name].data_ptr(), M0, 64, out_f32=ws.x0_f32.data_ptr(), out_bf16=ws.x0_bf.data_ptr())
        xf, xb = ws.x0_f32, ws.x0_bf
        for blk in self.blocks:
            c1, c2, cd, nm = blk["conv1"], blk["conv2"], blk["down"], blk["name"]
            M = B * c2.Ho * c2.Ho
            self._conv_fwd(c1, xb, B, ws)
            self._bn(c1, B, params, rstat, ws, train, update_running)
            self.bn_act(ws.y[c1.name].data_ptr(), ws.aff[c1.name].data_ptr(), M, c1.cout, out_bf16=ws.a1[nm].data_ptr())
            self._conv_fwd(c2, ws.a1[nm], B, ws)
            self._bn(c2, B, params, rstat, ws, train, update_running)
            if cd is not None:
                self._conv_fwd(cd, xb, B, ws)
                self._bn(cd, B, params, rstat, ws, train, update_running)
                self.bn_act(ws.y[c2.name].data_ptr(), ws.aff[c2.name].data_ptr(), M, c2.cout, res=ws.y[cd.name].data_ptr(), res_aff=ws.aff[cd.name].data_ptr(),
                            out_f32=ws.out_f32[nm].data_ptr(), out_bf16=ws.out_bf[nm].data_ptr())
            else:
                self.bn_act(ws.y[c2.name].data_ptr(), ws.aff[c2.name].data_ptr(), M, c2.cout, res=xf.data_ptr(), out_f32=ws.out_f32[nm].data_ptr(),
                            out_bf16=ws.out_bf[nm].data_ptr())
            xf, xb = ws.out_f32[nm], ws.out_bf[nm]

    def pool_forward(self, batch: int, ws=None):
        ws = self.ws if ws is None else ws
        last = self.blocks[-1]
        hw = last["conv2"].Ho ** 2
        check(self.lib.lc_nn_avgpool_forward(ws.out_f32[last["name"]].data_ptr(), batch, hw, 512, ws.feat.data_ptr(), stream_ptr()), "lc_nn_avgpool_forward")
        self.launches += 1

    def head_forward(self, batch: int, ncls: int, params=None, ws=None, logits=None):
        p = self.params if params is None else params
        ws = self.ws if ws is None else ws
        lg = self.logits if logits is None else logits
        self.pool_forward(batch, ws)
        check(self.lib.lc_linear_head(ws.feat.data_ptr(), _p(p, self.off_fc_w), _p(p, self.off_fc_b), batch, ncls, 512, lg.data_ptr(), self.cap, stream_ptr()),
              "lc_linear_head")
        self.launches += 1

    def loss(self, y: torch.Tensor, batch: int, ce_lo: int, ce_hi: int, pred_n: int, teacher_logits=None, kd_n: int = 0, kd_w: float = 0.0, T: float = 2.0):
        assert y.is_cuda and y.dtype == torch.int64
        check(self.lib.lc_loss_ce_kd(self.logits.data_ptr(), self.cap, _p(teacher_logits), self.cap, y.data_ptr(), batch, ce_lo, ce_hi, kd_n, kd_w, T, pred_n,
                                     self.dlogits.data_ptr(), self.pred.data_ptr(), self.scal.data_ptr(), stream_ptr()), "lc_loss_ce_kd")
        self.launches += 1

    def head_backward(self, batch: int, ncls: int):
        check(self.lib.lc_linear_head_backward(self.dlogits.data_ptr(), self.cap, self.ws.feat.data_ptr(), _p(self.params, self.off_fc_w), ncls, batch, 512,
                                               _p(self.grads, self.off_fc_w), _p(self.grads, self.off_fc_b), self.dfeat.data_ptr(), stream_ptr()),
              "lc_linear_head_backward")
        self.launches += 1
        self.pool_backward(self.dfeat[:batch])

    def pool_backward(self, gfeat: torch.Tensor):
        """d(features) [B][512] -> gradient of the last block's output (G[0], what `backward` starts from)."""
        assert gfeat.is_cuda and gfeat.dtype == torch.float32 and gfeat.is_contiguous()
        hw = self.blocks[-1]["conv2"].Ho ** 2
        check(self.lib.lc_nn_avgpool_backward(gfeat.data_ptr(), gfeat.shape[0], hw, 512, self.G[0].data_ptr(), stream_ptr()), "lc_nn_avgpool_backward")
        self.launches += 1

    # ---- backward ------------------------------------------------------------------------------------------------------------------------------
    def _wgrad(self, c: _Conv, B: int, colT_ptr: int, ldT: int):
        """dW = dY^T col: both operands M-contiguous ([cout][M] from a transpose of dy_bf, [K][M] patch matrix), split-K over the CTAs, fixed-order reduce."""
        M = B * c.Ho * c.Ho
        Mp = _up(M, 8)
        self.transpose_bf16(self.dy_bf.data_ptr(), c.cout, M, c.cout, self.dyT.data_ptr(), Mp)
        ns = self.wgrad_splits(c.cout, c.Kp, M)
        # contraction length Mp: the zero tails of both transposed operands make the padding exact
        self.gemm(self.dyT.data_ptr(), Mp, colT_ptr, ldT, self.wpart.data_ptr(), c.ldp, c.cout, c.Kp, Mp, ksplit=ns, stride_split=c.cout * c.ldp)
        self.wgrad_reduce(self.wpart.data_ptr(), ns, c.cout, c.cin, c.ks, KORDER_TAP_C, c.ldp, _p(self.grads, c.w_off))

    def _wgrad_from_act(self, c: _Conv, B: int, x_bf: torch.Tensor):
        M = B * c.Ho * c.Ho
        Mp = _up(M, 8)
        self.im2col(x_bf.data_ptr(), SRC_NHWC_BF16, B, c.H, c.H, c.cin, c.ks, c.stride, c.pad, KORDER_TAP_C, None, 0, self.colT.data_ptr(), Mp, c.Kp)
        self._wgrad(c, B, self.colT.data_ptr(), Mp)

    def _dgrad(self, c: _Conv, B: int, out: torch.Tensor, addend: Optional[torch.Tensor]):
        """d(input) of conv c from dy_bf: stride-1 3x3 as an implicit convolution with the flipped filter (addend through the residual epilogue),
        everything else through dcol = dY W and a gather col2im."""
        M = B * c.Ho * c.Ho
        if c.wd_mode == 2:
            self.conv(self.dy_bf.data_ptr(), c.wd.data_ptr(), out.data_ptr(), B, c.Ho, c.Ho, c.cout, c.cin, 3, 1, 1, residual=_p(addend))
        else:
            self.gemm(self.dy_bf.data_ptr(), c.cout, c.wd.data_ptr(), c.cout, self.dcol.data_ptr(), c.K, M, c.K, c.cout, out_f32=False)
            self.col2im(self.dcol.data_ptr(), c.K, _p(addend), out.data_ptr(), B, c.H, c.H, c.cin, c.ks, c.stride, c.pad, KORDER_TAP_C)

    def backward(self, x: torch.Tensor):
        """From G[0] = d(loss)/d(last block output) (set by head_backward / pool_backward) to every parameter gradient."""
        B = x.shape[0]
        ws, g = self.ws, self.grads
        self._pack_weights_bwd(self.params)
        cur = 0
        for bi in range(len(self.blocks) - 1, -1, -1):
            blk = self.blocks[bi]
            c1, c2, cd, nm = blk["conv1"], blk["conv2"], blk["down"], blk["name"]
            M = B * c2.Ho * c2.Ho
            x_bf = ws.x0_bf if bi == 0 else ws.out_bf[self.blocks[bi - 1]["name"]]
            G, Gn = self.G[cur], self.G[cur ^ 1]
            # bn2 (+ the ReLU of the block output): dy2, and the masked gradient that also flows into the shortcut
            self.bn_bwd(G.data_ptr(), ws.y[c2.name].data_ptr(), ws.aff[c2.name].data_ptr(), M, c2.cout, act_f32=ws.out_f32[nm].data_ptr(),
                        dgamma=_p(g, c2.g_off), dbeta=_p(g, c2.b_off), dy_bf16=self.dy_bf.data_ptr(), dz_out=self.Gz.data_ptr())
            self._wgrad_from_act(c2, B, ws.a1[nm])
            self._dgrad(c2, B, self.dA, None)
            if cd is not None:
                # shortcut: BN (no ReLU of its own) + 1x1 stride-2 conv
                self.bn_bwd(self.Gz.data_ptr(), ws.y[cd.name].data_ptr(), ws.aff[cd.name].data_ptr(), M, cd.cout, dgamma=_p(g, cd.g_off), dbeta=_p(g, cd.b_off),
                            dy_bf16=self.dy_bf.data_ptr())
                self._wgrad_from_act(cd, B, x_bf)
                self._dgrad(cd, B, self.dXd, None)
            # bn1 + ReLU: dy1
            self.bn_bwd(self.dA.data_ptr(), ws.y[c1.name].data_ptr(), ws.aff[c1.name].data_ptr(), M, c1.cout, act_bf16=ws.a1[nm].data_ptr(),
                        dgamma=_p(g, c1.g_off), dbeta=_p(g, c1.b_off), dy_bf16=self.dy_bf.data_ptr())
            self._wgrad_from_act(c1, B, x_bf)
            self._dgrad(c1, B, Gn, self.dXd if cd is not None else self.Gz)
            cur ^= 1
        # stem
        s = self.convs[0]
        M0 = B * s.Ho * s.Ho
        G = self.G[cur]
        if self.maxpool:
            check(self.lib.lc_nn_maxpool_backward(G.data_ptr(), ws.pool_idx.data_ptr(), B, self.img, self.img, 64, 3, 2, 1, self.dA0.data_ptr(), stream_ptr()),
                  "lc_nn_maxpool_backward")
            self.launches += 1
            self.bn_bwd(self.dA0.data_ptr(), ws.y[s.name].data_ptr(), ws.aff[s.name].data_ptr(), M0, 64, act_f32=ws.a0.data_ptr(), dgamma=_p(g, s.g_off),
                        dbeta=_p(g, s.b_off), dy_bf16=self.dy_bf.data_ptr())
        else:
            self.bn_bwd(G.data_ptr(), ws.y[s.name].data_ptr(), ws.aff[s.name].data_ptr(), M0, 64, act_f32=ws.x0_f32.data_ptr(), dgamma=_p(g, s.g_off),
                        dbeta=_p(g, s.b_off), dy_bf16=self.dy_bf.data_ptr())
        self._wgrad(s, B, ws.colT0.data_ptr(), _up(M0, 8))

    # ---- flat-arena kernels (same contracts as ResNetEngine) ---------------------------------------------------------------------------------------
    def ewc_penalty(self, theta_ref: torch.Tensor, fisher: torch.Tensor, lamda: float):
        if self._lamda_dev != lamda:
            self.hp[8] = lamda
            self._lamda_dev = lamda
        check(self.lib.lc_ewc_penalty_grad(self.params.data_ptr(), theta_ref.data_ptr(), fisher.data_ptr(), self.grads.data_ptr(), self.n_total,
                                           self.hp.data_ptr() + 32, self.flat_scratch.data_ptr(), self.flat_counter.data_ptr(), self.scal.data_ptr(), stream_ptr()),
              "lc_ewc_penalty_grad")
        self.launches += 1

    def fisher_accumulate(self, fisher: torch.Tensor, weight: float):
        check(self.lib.lc_fisher_accumulate(fisher.data_ptr(), self.grads.data_ptr(), self.n_total, float(weight), stream_ptr()), "lc_fisher_accumulate")
        self.launches += 1

    def sgd_step(self, momentum_buf: torch.Tensor, hp: torch.Tensor, grads: Optional[torch.Tensor] = None, frozen: Optional[Tuple[int, int]] = None):
        g = self.grads if grads is None else grads
        if frozen is not None:
            check(self.lib.lc_sgd_momentum_frozen(self.params.data_ptr(), g.data_ptr(), momentum_buf.data_ptr(), self.n_total, hp.data_ptr(), frozen[0], frozen[1],
                                                  stream_ptr()), "lc_sgd_momentum_frozen")
        else:
            check(self.lib.lc_sgd_momentum(self.params.data_ptr(), g.data_ptr(), momentum_buf.data_ptr(), self.n_total, hp.data_ptr(), stream_ptr()),
                  "lc_sgd_momentum")
        self.launches += 1

    def new_teacher_workspace(self):
        return _R18Workspace(self, self.max_batch)

    def teacher_logits(self, t, x: torch.Tensor) -> torch.Tensor:
        B = x.shape[0]
        self.forward(x, train=False, update_running=False, params=t.params, rstat=t.rstat, ws=t.ws)
        self.head_forward(B, t.ncls, params=t.params, ws=t.ws, logits=t.logits)
        return t.logits


# ======================================================================================================================================================
# AlexNet_TRGP
# ======================================================================================================================================================
ALEXNET_SPEC = (  # name, bn, cin, cout, ks, H_in, H_out (conv) / in, out (linear), dropout p, pooled H
    ("conv1", "bn1", 3, 64, 4, 32, 29, 0.2, 14), ("conv2", "bn2", 64, 128, 3, 14, 12, 0.2, 6), ("conv3", "bn3", 128, 256, 2, 6, 5, 0.5, 2),
    ("fc1", "bn4", 1024, 2048, 1, 1, 1, 0.5, 1), ("fc2", "bn5", 2048, 2048, 1, 1, 1, 0.5, 1))


def alexnet_param_layout():
    """(name, shape) in the reference's registration order (alexnet.py:100-114)."""
    out = []
    for name, bn, cin, cout, ks, *_ in ALEXNET_SPEC:
        out += [(name + ".weight", (cout, cin, ks, ks) if name.startswith("conv") else (cout, cin)), (bn + ".weight", (cout,)), (bn + ".bias", (cout,))]
    return out


class _ALayer:
    __slots__ = ("name", "bn", "cin", "cout", "ks", "H", "Ho", "p", "Hp", "K", "w_off", "g_off", "b_off", "wk", "wT", "col", "colT", "y", "aff", "a", "pool_f32",
                 "pool_bf", "idx", "M", "Mp", "nsplit")


class AlexNetEngine(NNOps):
    """`AlexNet_TRGP.forward` (alexnet.py:124-156) + backward.  Every layer is patch matrix -> tcgen05 GEMM (K in weight.view(out, -1) order, the order
    of GPM's bases) -> BatchNorm with batch statistics (track_running_stats=False: also in eval) -> ReLU -> dropout -> MaxPool2d(2).  The trainable
    arena `theta` holds the backbone parameters in registration order followed by the per-task bias-free heads (gpm.py:29-32)."""

    def __init__(self, max_batch: int = 128, head_sizes=(10,), device=None):
        super().__init__(device)
        dev = self.device
        self.dev = dev
        self.max_batch, self.feat_dim, self.img, self.in_ch = max_batch, 2048, 32, 3
        self.layout = alexnet_param_layout()
        self.param_off: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape in self.layout:
            n = 1
            for d in shape:
                n *= d
            self.param_off[name] = (off, shape)
            off += n
        self.n_backbone = off
        self.head_sizes = list(head_sizes)
        self.head_off = []
        for n in self.head_sizes:
            self.head_off.append(off)
            off += n * self.feat_dim
        self.n_total = _up(off, 4)
        self.theta = torch.zeros(self.n_total, device=dev)
        self.theta_grad = torch.zeros(self.n_total, device=dev)
        B = max_batch
        f32, bf = torch.float32, torch.bfloat16
        e = lambda *s, dt=f32: torch.zeros(*s, device=dev, dtype=dt)
        self.layers: List[_ALayer] = []
        for (name, bn, cin, cout, ks, H, Ho, p, Hp) in ALEXNET_SPEC:
            L = _ALayer()
            L.name, L.bn, L.cin, L.cout, L.ks, L.H, L.Ho, L.p, L.Hp = name, bn, cin, cout, ks, H, Ho, p, Hp
            L.K = cin * ks * ks
            L.M = B * Ho * Ho
            L.Mp = _up(L.M, 8)
            L.w_off, L.g_off, L.b_off = self.param_off[name + ".weight"][0], self.param_off[bn + ".weight"][0], self.param_off[bn + ".bias"][0]
            L.wk, L.wT = e(cout, L.K, dt=bf), e(L.K, cout, dt=bf)
            L.col, L.colT = e(L.M, L.K, dt=bf), e(L.K, L.Mp, dt=bf)
            L.y, L.aff, L.a = e(L.M, cout), e(4 * cout), e(L.M, cout)
            if name.startswith("conv"):
                L.pool_f32, L.pool_bf = e(B * Hp * Hp, cout), e(B * Hp * Hp, cout, dt=bf)
                L.idx = torch.zeros(B * Hp * Hp, cout, device=dev, dtype=torch.uint8)
            L.nsplit = self.wgrad_splits(cout, L.K, L.M)
            self.layers.append(L)
        self.a_bf = [e(B, 2048, dt=bf) for _ in range(2)]          # BF16 copies of the linear layers' outputs (next GEMM operand)
        gmax = max(L.M * L.cout for L in self.layers)
        self.G = [e(gmax), e(gmax)]
        self.dy_bf = e(gmax, dt=bf)
        self.dyT = e(max(L.cout * L.Mp for L in self.layers), dt=bf)
        self.dcol = e(max(L.M * L.K for L in self.layers[1:]), dt=bf)
        self.wpart = e(max(L.nsplit * L.cout * _up(L.K, 4) for L in self.layers))
        self.dpool = e(max(B * L.Hp * L.Hp * L.cout for L in self.layers[:3]))
        ncls = max(self.head_sizes)
        self.cap = ncls
        self.logits_all = e(B, sum(self.head_sizes))
        self.logits = e(B, ncls)
        self.dlogits = e(B, ncls)
        self.pred = torch.zeros(B, dtype=torch.int64, device=dev)
        self.scal = e(8)
        self.dfeat = e(B, 2048)
        self.rng = torch.tensor([0x5EED, 0], dtype=torch.int64, device=dev)      # dropout: {seed, step}
        self.yshift = torch.zeros(B, dtype=torch.int64, device=dev)

    # ---- views -----------------------------------------------------------------------------------------------------------------------------------
    def param_view(self, name: str, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        off, shape = self.param_off[name]
        n = 1
        for d in shape:
            n *= d
        return (self.theta if arena is None else arena)[off:off + n].view(shape)

    def head_view(self, t: int, arena: Optional[torch.Tensor] = None) -> torch.Tensor:
        a = self.theta if arena is None else arena
        return a[self.head_off[t]:self.head_off[t] + self.head_sizes[t] * self.feat_dim].view(self.head_sizes[t], self.feat_dim)

    def features(self, batch: int) -> torch.Tensor:
        return self.layers[-1].a[:batch]

    # ---- forward ---------------------------------------------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, train: bool, need_backward: Optional[bool] = None):
        """train: dropout on (masks from the device-side {seed, step} pair).  BatchNorm uses batch statistics in either mode."""
        B = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape[1:]) == (3, 32, 32) and B <= self.max_batch
        need_backward = train if need_backward is None else need_backward
        th = self.theta
        src, kind, Hin, Cin = x.data_ptr(), SRC_NCHW_F32, 32, 3
        for i, L in enumerate(self.layers):
            M = B * L.Ho * L.Ho
            self.pack(_p(th, L.w_off), L.cout, L.cin, L.ks, KORDER_C_TAP, 0, L.wk.data_ptr(), L.K)
            if i == 3:                       # x.view(B, -1) of the NCHW tensor (alexnet.py:144): a 2x2 "patch" in (c, kh, kw) order is exactly that flatten
                self.im2col(src, kind, B, 2, 2, 256, 2, 1, 0, KORDER_C_TAP, L.col.data_ptr(), L.K, L.colT.data_ptr() if need_backward else None, _up(M, 8), L.K)
            elif i == 4:
                self.cast_transpose(self.layers[3].a.data_ptr(), B, 2048, L.col.data_ptr(), L.K, L.colT.data_ptr() if need_backward else None, _up(M, 8))
            else:
                self.im2col(src, kind, B, Hin, Hin, Cin, L.ks, 1, 0, KORDER_C_TAP, L.col.data_ptr(), L.K, L.colT.data_ptr() if need_backward else None, _up(M, 8), L.K)
            self.gemm(L.col.data_ptr(), L.K, L.wk.data_ptr(), L.K, L.y.data_ptr(), L.cout, M, L.cout, L.K)
            self.bn_stats(L.y.data_ptr(), M, L.cout, _p(th, L.g_off), _p(th, L.b_off), None, L.aff.data_ptr())
            self.bn_act(L.y.data_ptr(), L.aff.data_ptr(), M, L.cout, drop_p=L.p if train else 0.0, rng=self.rng.data_ptr() if train else None, rng_stream=i,
                        out_f32=L.a.data_ptr())
            if i < 3:
                check(self.lib.lc_nn_maxpool_forward(L.a.data_ptr(), B, L.Ho, L.Ho, L.cout, 2, 2, 0, L.pool_f32.data_ptr(), L.pool_bf.data_ptr(), L.idx.data_ptr(),
                                                     stream_ptr()), "lc_nn_maxpool_forward")
                self.launches += 1
                src, kind, Hin, Cin = L.pool_bf.data_ptr(), SRC_NHWC_BF16, L.Hp, L.cout

    def heads_forward(self, batch: int, task: Optional[int] = None):
        """logits of one head into `logits` (training), or of every head side by side into `logits_all` (gpm.py:34-41)."""
        feat = self.layers[-1].a
        if task is not None:
            check(self.lib.lc_linear_head(feat.data_ptr(), _p(self.theta, self.head_off[task]), None, batch, self.head_sizes[task], 2048, self.logits.data_ptr(),
                                          self.cap, stream_ptr()), "lc_linear_head")
            self.launches += 1
            return self.logits[:batch, :self.head_sizes[task]]
        tot = sum(self.head_sizes)
        check(self.lib.lc_linear_head(feat.data_ptr(), _p(self.theta, self.head_off[0]), None, batch, tot, 2048, self.logits_all.data_ptr(), tot, stream_ptr()),
              "lc_linear_head")                   # the heads are contiguous rows of one [total][2048] matrix
        self.launches += 1
        return self.logits_all[:batch]

    def loss_backward_head(self, y_shifted: torch.Tensor, batch: int, task: int):
        n = self.head_sizes[task]
        check(self.lib.lc_loss_ce_kd(self.logits.data_ptr(), self.cap, None, self.cap, y_shifted.data_ptr(), batch, 0, n, 0, 0.0, 2.0, n, self.dlogits.data_ptr(),
                                     self.pred.data_ptr(), self.scal.data_ptr(), stream_ptr()), "lc_loss_ce_kd")
        check(self.lib.lc_linear_head_backward(self.dlogits.data_ptr(), self.cap, self.layers[-1].a.data_ptr(), _p(self.theta, self.head_off[task]), n, batch, 2048,
                                               _p(self.theta_grad, self.head_off[task]), None, self.dfeat.data_ptr(), stream_ptr()), "lc_linear_head_backward")
        self.launches += 2

    # ---- backward --------------------------------------------------------------------------------------------------------------------------------
    def backward(self, batch: int):
        """From `dfeat` = d(loss)/d(features) to the gradient of every backbone parameter (in `theta_grad`)."""
        B, th, g = batch, self.theta, self.theta_grad
        gin = self.dfeat
        for i in range(4, -1, -1):
            L = self.layers[i]
            M = B * L.Ho * L.Ho
            Mp = _up(M, 8)
            if i < 3:        # gradient arrives at the pooled map: route it back to the window maxima
                check(self.lib.lc_nn_maxpool_backward(gin.data_ptr(), L.idx.data_ptr(), B, L.Ho, L.Ho, L.cout, 2, 2, 0, self.G[0].data_ptr(), stream_ptr()),
                      "lc_nn_maxpool_backward")
                self.launches += 1
                gin = self.G[0]
            # BN + ReLU + dropout (a kept unit carries 1/(1-p), a dropped one is stored as 0: one mask)
            self.bn_bwd(gin.data_ptr(), L.y.data_ptr(), L.aff.data_ptr(), M, L.cout, act_f32=L.a.data_ptr(), gscale=1.0 / (1.0 - L.p) if self._train_step else 1.0,
                        dgamma=_p(g, L.g_off), dbeta=_p(g, L.b_off), dy_bf16=self.dy_bf.data_ptr())
            # dW = dY^T col
            self.transpose_bf16(self.dy_bf.data_ptr(), L.cout, M, L.cout, self.dyT.data_ptr(), Mp)
            ns, ldp = self.wgrad_splits(L.cout, L.K, M), _up(L.K, 4)
            if ns > 1:
                self.gemm(self.dyT.data_ptr(), Mp, L.colT.data_ptr(), Mp, self.wpart.data_ptr(), ldp, L.cout, L.K, Mp, ksplit=ns, stride_split=L.cout * ldp)
                self.wgrad_reduce(self.wpart.data_ptr(), ns, L.cout, L.cin, L.ks, KORDER_C_TAP, ldp, _p(g, L.w_off))
            else:
                self.gemm(self.dyT.data_ptr(), Mp, L.colT.data_ptr(), Mp, _p(g, L.w_off), L.K, L.cout, L.K, Mp)
            if i == 0:
                break
            # d(input) = fold(dY W)
            self.pack(_p(th, L.w_off), L.cout, L.cin, L.ks, KORDER_C_TAP, 1, L.wT.data_ptr(), L.cout)
            P = self.layers[i - 1]
            if i == 4:
                self.gemm(self.dy_bf.data_ptr(), L.cout, L.wT.data_ptr(), L.cout, self.G[1].data_ptr(), L.K, M, L.K, L.cout)      # fp32 d(a4) directly
                gin = self.G[1]
            else:
                self.gemm(self.dy_bf.data_ptr(), L.cout, L.wT.data_ptr(), L.cout, self.dcol.data_ptr(), L.K, M, L.K, L.cout, out_f32=False)
                if i == 3:
                    self.col2im(self.dcol.data_ptr(), L.K, None, self.dpool.data_ptr(), B, 2, 2, 256, 2, 1, 0, KORDER_C_TAP)
                else:
                    self.col2im(self.dcol.data_ptr(), L.K, None, self.dpool.data_ptr(), B, P.Hp, P.Hp, P.cout, L.ks, 1, 0, KORDER_C_TAP)
                gin = self.dpool

    _train_step = True

    # ---- GPM's representation matrices (gpm.py:144-168) on the device ------------------------------------------------------------------------------------
    def representation_matrices(self, x125: torch.Tensor, batch_list=(24, 100, 100)) -> List[torch.Tensor]:
        """Eval-mode forward of the 125 selected samples, then per TRGP layer the fp32 matrix whose columns are the layer's input patches:
        conv: [Cin*k*k][n*s*s] of the first 24 / 100 / 100 samples (the reference's Python triple loop, here one im2col launch each); linear: input^T."""
        B = x125.shape[0]
        self.forward(x125, train=False, need_backward=False)
        mats = []
        srcs = [(x125.data_ptr(), SRC_NCHW_F32, 32, 3)] + [(L.pool_f32.data_ptr(), SRC_NHWC_F32, L.Hp, L.cout) for L in self.layers[:2]]
        for (ptr, kind, H, C), L, n in zip(srcs, self.layers[:3], batch_list):
            n = min(n, B)
            M = n * L.Ho * L.Ho
            m = torch.empty(L.K, M, device=self.device)
            check(self.lib.lc_nn_im2col_f32(ptr, kind, n, H, H, C, L.ks, 1, 0, KORDER_C_TAP, None, 0, m.data_ptr(), M, stream_ptr()), "lc_nn_im2col_f32")
            mats.append(m)
        f0 = torch.empty(1024, B, device=self.device)
        check(self.lib.lc_nn_im2col_f32(self.layers[2].pool_f32.data_ptr(), SRC_NHWC_F32, B, 2, 2, 256, 2, 1, 0, KORDER_C_TAP, None, 0, f0.data_ptr(), B, stream_ptr()),
              "lc_nn_im2col_f32")
        mats.append(f0)
        mats.append(self.layers[3].a[:B].t().contiguous())
        self.launches += 4
        return mats
