"""Device-side state and kernel sequencing for the CIFAR-ResNet hot path.

`ResNetEngine` owns the flat fp32 arenas (parameters, gradients, BN running statistics), the opaque activation workspace
and the C plan handle, and exposes the per-step building blocks that the method classes in `libcontinual_b200.model`
compose: backbone forward / backward, head + loss, EWC penalty, Fisher accumulation.  Everything here is launch plumbing:
torch supplies device memory and streams, every FLOP runs in `liblc_b200.so`.

HBM layout (all fp32, one allocation each):
  params : [ backbone parameters in the reference's named_parameters() order | fc.weight (cap x feat) | fc.bias (cap) ]
  grads  : same layout (what autograd would have put in .grad)
  rstat  : per BN layer (running_mean[C], running_var[C])
  ws     : plan workspace (activations NHWC, packed weights, partial sums, scratch); zero-filled once
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib
from ._lib import check, ptr, stream_ptr

LC_WS_FEAT, LC_WS_GRAD_LAST, LC_WS_FMAP1, LC_WS_FMAP2, LC_WS_FMAP3, LC_WS_DFEAT = range(6)
FLAT_SCRATCH_FLOATS = 2 * 296 + 8


def _round4(v: int) -> int:
    return (v + 3) // 4 * 4


def cifar_resnet_param_layout(depth: int = 32, in_ch: int = 3, style: str = "cifar") -> Tuple[List[Tuple[str, Tuple[int, ...]]], List[str]]:
    """(name, shape) of every backbone parameter in module registration order, and the BN layer prefixes in buffer order.
    style 'cifar': `CifarResNet` (core/model/backbone/resnet.py:334-340,294-301,361-376);
    style 'lucir': `modified_ResNet` / `modified_BasicBlock` (resnet.py:472-547) — same topology and registration order, other names."""
    nblk = (depth - 2) // 6
    lucir = style == "lucir"
    stem_c, stem_bn = ("conv1", "bn1") if lucir else ("conv_1_3x3", "bn_1")
    ca, ba, cb, bb = ("conv1", "bn1", "conv2", "bn2") if lucir else ("conv_a", "bn_a", "conv_b", "bn_b")
    params: List[Tuple[str, Tuple[int, ...]]] = [(stem_c + ".weight", (16, in_ch, 3, 3)), (stem_bn + ".weight", (16,)), (stem_bn + ".bias", (16,))]
    bns = [stem_bn]
    inpl = 16
    for s, planes in enumerate((16, 32, 64), start=1):
        for b in range(nblk):
            pre = f"layer{s}.{b}" if lucir else f"stage_{s}.{b}"
            cin = inpl if b == 0 else planes
            params += [(pre + f".{ca}.weight", (planes, cin, 3, 3)), (pre + f".{ba}.weight", (planes,)), (pre + f".{ba}.bias", (planes,)),
                       (pre + f".{cb}.weight", (planes, planes, 3, 3)), (pre + f".{bb}.weight", (planes,)), (pre + f".{bb}.bias", (planes,))]
            bns += [pre + "." + ba, pre + "." + bb]
            if b == 0 and s > 1:
                params += [(pre + ".downsample.0.weight", (planes, cin, 1, 1)), (pre + ".downsample.1.weight", (planes,)),
                           (pre + ".downsample.1.bias", (planes,))]
                bns.append(pre + ".downsample.1")
        inpl = planes
    return params, bns


class TeacherState:
    """Frozen copy of a network (parameters + running statistics) with its own workspace: the KD teacher of
    iCaRL (`old_network`, icarl.py:172-173) and LwF (`old_backbone`/`old_fc`, lwf.py:28-50)."""

    def __init__(self, eng: "ResNetEngine"):
        self.params = eng.params.clone()
        self.rstat = eng.rstat.clone()
        self.ws = eng.new_teacher_workspace() if hasattr(eng, "new_teacher_workspace") else torch.zeros_like(eng.ws)
        self.logits = torch.zeros_like(eng.logits)
        self.ncls = eng.ncls


class ResNetEngine:
    def __init__(self, depth: int = 32, max_batch: int = 128, num_class_cap: int = 100, device=None, in_ch: int = 3, img: int = 32,
                 style: str = "cifar", last_relu: bool = True):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise _lib.LcError("libcontinual_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            check(self.lib.lc_device_check(), "lc_device_check: built for sm_100a only")
            h = ctypes.c_void_p()
            check(self.lib.lc_resnet_create(depth, in_ch, img, max_batch, ctypes.byref(h)), "lc_resnet_create")
        self.h = h
        self.depth, self.max_batch, self.cap, self.feat_dim, self.img, self.in_ch = depth, max_batch, num_class_cap, 64, img, in_ch
        self.n_backbone = int(self.lib.lc_resnet_param_count(h))
        self.off_fc_w = _round4(self.n_backbone)
        self.off_fc_b = self.off_fc_w + self.cap * self.feat_dim
        self.n_total = _round4(self.off_fc_b + self.cap)
        dev = self.device
        self.params = torch.zeros(self.n_total, device=dev)
        self.grads = torch.zeros(self.n_total, device=dev)
        self.rstat = torch.zeros(int(self.lib.lc_resnet_rstat_count(h)), device=dev)
        self.ws = torch.zeros(int(self.lib.lc_resnet_workspace_floats(h)), device=dev)
        self.logits = torch.zeros(max_batch, self.cap, device=dev)
        self.dlogits = torch.zeros(max_batch, self.cap, device=dev)
        self.pred = torch.zeros(max_batch, dtype=torch.int64, device=dev)
        self.scal = torch.zeros(8, device=dev)             # [0] loss [1] #correct [2] ce [3] kd [4] ewc penalty
        self.flat_scratch = torch.zeros(FLAT_SCRATCH_FLOATS, device=dev)
        self.flat_counter = torch.zeros(4, dtype=torch.int32, device=dev)
        self.hp = torch.zeros(16, device=dev)              # [0..2] sgd lr/mu/wd  [8] ewc lamda
        self._lamda_dev = None
        self.autograd_grads = None                         # per-step copy of `grads` handed to autograd (p.grad are views of it)
        self.ncls = 0                                      # live rows of the head
        self.layout, self.bn_names = cifar_resnet_param_layout(depth, in_ch, style)
        check(self.lib.lc_resnet_set_last_relu(h, int(last_relu)), "lc_resnet_set_last_relu")
        self.cos_inv_norm = torch.zeros(max_batch + self.cap, device=dev)      # cosine head: 1/||feat_b||, 1/||w_c||
        self.scores = torch.zeros(max_batch, self.cap, device=dev)             # cosine head: scores before the sigma scale
        self.dscores = torch.zeros(max_batch, self.cap, device=dev)
        self.dfeat_extra = torch.zeros(max_batch, self.feat_dim, device=dev)   # loss terms acting directly on the features
        # offsets of every named parameter and BN layer, cross-checked against the C plan
        self.param_off: Dict[str, Tuple[int, Tuple[int, ...]]] = {}
        off = 0
        for name, shape in self.layout:
            n = 1
            for d in shape:
                n *= d
            self.param_off[name] = (off, shape)
            off += n
        assert off == self.n_backbone, (off, self.n_backbone)
        self.rstat_off: Dict[str, Tuple[int, int]] = {}
        nconv = int(self.lib.lc_resnet_num_convs(h))
        assert nconv == len(self.bn_names)
        conv_names = [n for n, s in self.layout if len(s) == 4]
        for i in range(nconv):
            w_off, cout, ro, go = ctypes.c_longlong(), ctypes.c_int(), ctypes.c_longlong(), ctypes.c_longlong()
            check(self.lib.lc_resnet_conv_info(h, i, ctypes.byref(w_off), ctypes.byref(cout), None, None, None, ctypes.byref(go), None, ctypes.byref(ro)))
            assert self.param_off[conv_names[i]][0] == w_off.value and self.param_off[self.bn_names[i] + ".weight"][0] == go.value
            self.rstat_off[self.bn_names[i]] = (ro.value, cout.value)
        self.reset_running_stats()
        self._off = {k: int(self.lib.lc_resnet_ws_offset(h, k)) for k in range(6)}
        self.launches = 0
        self.precision = "fp32"

    def set_precision(self, precision: str):
        """'fp32': exact CUDA-core convolutions; 'tc': tcgen05 tensor-core convolutions for the stride-1 3x3 layers (TF32 operands
        for forward / data gradient, BF16 operands for the weight gradient, fp32 accumulation in TMEM)."""
        alias = {"fp32": "fp32", "exact": "fp32", "tc": "tc", "tf32": "tc"}
        if precision not in alias:
            raise ValueError(f"precision must be one of {sorted(alias)} ('tf32' is an alias of 'tc'), got {precision!r}")
        precision = alias[precision]
        mode = {"fp32": 0, "tc": 1}[precision]
        check(self.lib.lc_resnet_set_mode(self.h, mode), "lc_resnet_set_mode")
        self.precision = precision

    def tensor_core_error(self) -> bool:
        """True if any tensor-core completion barrier ever timed out (results would be garbage)."""
        return bool(int(self.ws[:16].view(torch.int32)[8]) != 0)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.lc_resnet_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- views ------------------------------------------------------------------------------------------------------
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

    def features(self, batch: int, ws: Optional[torch.Tensor] = None) -> torch.Tensor:
        w = self.ws if ws is None else ws
        o = self._off[LC_WS_FEAT]
        return w[o:o + batch * self.feat_dim].view(batch, self.feat_dim)

    def fmaps(self, batch: int):
        """Stage outputs as NCHW-shaped views of the NHWC workspace tensors (torch channels_last memory)."""
        outs = []
        for k, (c, s) in zip((LC_WS_FMAP1, LC_WS_FMAP2, LC_WS_FMAP3), ((16, 32), (32, 16), (64, 8))):
            o = self._off[k]
            outs.append(self.ws[o:o + batch * s * s * c].view(batch, s, s, c).permute(0, 3, 1, 2))
        return outs

    # ---- kernels sequences ---------------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor, train: bool, update_running: bool = True, params=None, rstat=None, ws=None):
        B = x.shape[0]
        assert x.is_cuda and x.dtype == torch.float32 and x.is_contiguous() and tuple(x.shape[1:]) == (self.in_ch, self.img, self.img)
        check(self.lib.lc_resnet_forward(self.h, ptr(x), B, ptr(self.params if params is None else params), ptr(self.rstat if rstat is None else rstat),
                                         ptr(self.ws if ws is None else ws), int(train), int(update_running), stream_ptr()), "lc_resnet_forward")
        self.launches += int(self.lib.lc_resnet_num_launches(self.h, 0))

    def pool_forward(self, batch: int, ws=None):
        w = self.ws if ws is None else ws
        check(self.lib.lc_avgpool_forward(w.data_ptr() + 4 * self._off[LC_WS_FMAP3], batch, 64, self.feat_dim, w.data_ptr() + 4 * self._off[LC_WS_FEAT],
                                          stream_ptr()), "lc_avgpool_forward")
        self.launches += 1

    def pool_backward(self, gfeat: torch.Tensor):
        """d(features) [B][64] -> gradient of the last feature map (the slot lc_resnet_backward consumes)."""
        assert gfeat.is_cuda and gfeat.dtype == torch.float32 and gfeat.is_contiguous()
        check(self.lib.lc_avgpool_backward(ptr(gfeat), gfeat.shape[0], 64, self.feat_dim, self.ws.data_ptr() + 4 * self._off[LC_WS_GRAD_LAST], stream_ptr()),
              "lc_avgpool_backward")
        self.launches += 1

    def head_forward(self, batch: int, ncls: int, params=None, ws=None, logits=None):
        p = self.params if params is None else params
        w = self.ws if ws is None else ws
        lg = self.logits if logits is None else logits
        o = self._off[LC_WS_FMAP3]
        check(self.lib.lc_head_forward(w.data_ptr() + 4 * o, batch, 64, self.feat_dim, p.data_ptr() + 4 * self.off_fc_w, p.data_ptr() + 4 * self.off_fc_b,
                                       ncls, w.data_ptr() + 4 * self._off[LC_WS_FEAT], ptr(lg), self.cap, stream_ptr()), "lc_head_forward")
        self.launches += 1

    def loss(self, y: torch.Tensor, batch: int, ce_lo: int, ce_hi: int, pred_n: int, teacher_logits=None, kd_n: int = 0, kd_w: float = 0.0, T: float = 2.0):
        assert y.is_cuda and y.dtype == torch.int64
        check(self.lib.lc_loss_ce_kd(ptr(self.logits), self.cap, ptr(teacher_logits), self.cap, ptr(y), batch, ce_lo, ce_hi, kd_n, kd_w, T, pred_n,
                                     ptr(self.dlogits), ptr(self.pred), ptr(self.scal), stream_ptr()), "lc_loss_ce_kd")
        self.launches += 1

    def head_backward(self, batch: int, ncls: int):
        g, w = self.grads, self.ws
        check(self.lib.lc_head_backward(ptr(self.dlogits), self.cap, w.data_ptr() + 4 * self._off[LC_WS_FEAT], self.params.data_ptr() + 4 * self.off_fc_w,
                                        ncls, batch, self.feat_dim, g.data_ptr() + 4 * self.off_fc_w, g.data_ptr() + 4 * self.off_fc_b,
                                        w.data_ptr() + 4 * self._off[LC_WS_DFEAT], w.data_ptr() + 4 * self._off[LC_WS_GRAD_LAST], 64, stream_ptr()),
              "lc_head_backward")
        self.launches += 1

    def backward(self, x: torch.Tensor):
        check(self.lib.lc_resnet_backward(self.h, ptr(x), x.shape[0], ptr(self.params), ptr(self.ws), ptr(self.grads), stream_ptr()), "lc_resnet_backward")
        self.launches += int(self.lib.lc_resnet_num_launches(self.h, 1))

    def ewc_penalty(self, theta_ref: torch.Tensor, fisher: torch.Tensor, lamda: float):
        if self._lamda_dev != lamda:       # host mirror: no device read-back on the step path
            self.hp[8] = lamda
            self._lamda_dev = lamda
        check(self.lib.lc_ewc_penalty_grad(ptr(self.params), ptr(theta_ref), ptr(fisher), ptr(self.grads), self.n_total, self.hp.data_ptr() + 32,
                                           ptr(self.flat_scratch), ptr(self.flat_counter), ptr(self.scal), stream_ptr()), "lc_ewc_penalty_grad")
        self.launches += 1

    def fisher_accumulate(self, fisher: torch.Tensor, weight: float):
        check(self.lib.lc_fisher_accumulate(ptr(fisher), ptr(self.grads), self.n_total, float(weight), stream_ptr()), "lc_fisher_accumulate")
        self.launches += 1

    def sgd_step(self, momentum_buf: torch.Tensor, hp: torch.Tensor, grads: Optional[torch.Tensor] = None, frozen: Optional[Tuple[int, int]] = None):
        g = self.grads if grads is None else grads
        if frozen is not None:
            check(self.lib.lc_sgd_momentum_frozen(ptr(self.params), ptr(g), ptr(momentum_buf), self.n_total, ptr(hp), frozen[0], frozen[1], stream_ptr()),
                  "lc_sgd_momentum_frozen")
            self.launches += 1
            return
        check(self.lib.lc_sgd_momentum(ptr(self.params), ptr(g), ptr(momentum_buf), self.n_total, ptr(hp), stream_ptr()), "lc_sgd_momentum")
        self.launches += 1

    # ---- composite: logits of a frozen teacher on the same batch -------------------------------------------------------
    def teacher_logits(self, t: TeacherState, x: torch.Tensor) -> torch.Tensor:
        B = x.shape[0]
        self.forward(x, train=False, update_running=False, params=t.params, rstat=t.rstat, ws=t.ws)
        self.head_forward(B, t.ncls, params=t.params, ws=t.ws, logits=t.logits)
        return t.logits
