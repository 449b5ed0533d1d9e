"""Device-side state and kernel sequencing for the ViT-B/16 hot path (L2P / prompt methods: frozen backbone, trainable prompts + head).

`ViTEngine` owns the frozen backbone weights (fp32 master copy for the row-wise kernels, BF16 copies in both orientations for the
tcgen05 GEMMs), per-(batch, tokens) activation workspaces and the launch sequence of
`VisionTransformer.forward(prompt_flag='l2p')` (core/model/backbone/transformer.py:2222-2261):

    patchify -> patch-embed GEMM (+bias +pos_embed) -> [prompts | cls | patches]
    12 x { LN1 -> QKV GEMM -> fused attention (S, softmax, P V on chip) -> proj GEMM (+residual) -> LN2 -> fc1 GEMM (+GELU) -> fc2 GEMM (+residual) }
    final LN (eps 1e-6)

and of its backward with respect to the INPUT tokens only (the backbone is frozen: l2p.py:66-71), which is what carries the loss
gradient back to the prompt rows.  The residual stream is fp32; GEMM operands, the stored QKV / attention / GELU outputs are BF16;
accumulation is fp32 in TMEM.  Everything here is launch plumbing: torch supplies device memory and streams, every FLOP runs in
`liblc_b200.so`.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib
from ._lib import GemmDesc, check, stream_ptr

DIM, HEADS, HDIM, MLP, GRID, PATCHES = 768, 12, 64, 3072, 14, 196
COV_SPLIT = 8          # split-K factor of the input-matrix GEMM (18 output tiles x 8 = 144 CTAs)


def vit_param_layout(depth: int = 12):
    """(name, shape) in the reference `VisionTransformer` registration order (transformer.py:2186-2205, block: :1320-1326)."""
    out = [("cls_token", (1, 1, DIM)), ("pos_embed", (1, PATCHES + 1, DIM)), ("patch_embed.proj.weight", (DIM, 3, 16, 16)),
           ("patch_embed.proj.bias", (DIM,))]
    for i in range(depth):
        b = f"transformer.blocks.{i}."
        out += [(b + "attn.qkv.weight", (3 * DIM, DIM)), (b + "attn.qkv.bias", (3 * DIM,)), (b + "attn.proj.weight", (DIM, DIM)),
                (b + "attn.proj.bias", (DIM,)), (b + "ln_1.weight", (DIM,)), (b + "ln_1.bias", (DIM,)), (b + "mlp.fc1.weight", (MLP, DIM)),
                (b + "mlp.fc1.bias", (MLP,)), (b + "mlp.fc2.weight", (DIM, MLP)), (b + "mlp.fc2.bias", (DIM,)), (b + "ln_2.weight", (DIM,)),
                (b + "ln_2.bias", (DIM,))]
    out += [("norm.weight", (DIM,)), ("norm.bias", (DIM,))]
    return out


def _up8(v: int) -> int:
    return (v + 7) // 8 * 8


class _Workspace:
    """Activation buffers of one (batch, tokens) shape.  With `save` every block keeps what its backward needs:
    x_in / x_mid (LayerNorm inputs), qkv, attention output + row log-sum-exp, GELU'(fc1 output) (`upre`)."""

    def __init__(self, B: int, T: int, depth: int, save: bool, dev, lora_cols: int = 0, keep_h: bool = False):
        self.B, self.T, self.Tp, self.save = B, T, _up8(T), save
        bf, f32 = torch.bfloat16, torch.float32
        n = B * T
        nx = depth + 1 if save else 2
        self.x = [torch.empty(B, T, DIM, device=dev, dtype=f32) for _ in range(nx)]          # block inputs (x[depth] = last block's output)
        self.xmid = [torch.empty(B, T, DIM, device=dev, dtype=f32) for _ in range(depth if save else 1)]
        self.qkv = [torch.empty(n, 3 * DIM, device=dev, dtype=bf) for _ in range(depth if save else 1)]
        self.o = [torch.empty(n, DIM, device=dev, dtype=bf) for _ in range(depth if save else 1)]          # attention output (pre-projection)
        self.lse = [torch.empty(B, HEADS, T, device=dev, dtype=f32) for _ in range(depth if save else 1)]   # base-2 log-sum-exp of the score rows
        self.upre = [torch.empty(n, MLP, device=dev, dtype=bf) for _ in range(depth if save else 1)]
        self.patches = torch.empty(B * PATCHES, DIM, device=dev, dtype=bf)
        self.h = torch.empty(n, DIM, device=dev, dtype=bf)
        self.u = torch.empty(n, MLP, device=dev, dtype=bf)
        self.y = torch.empty(B, T, DIM, device=dev, dtype=f32)
        self.ystat = torch.empty(n, 2, device=dev, dtype=f32)
        self.feat = torch.empty(B, DIM, device=dev, dtype=f32)
        # LoRA down-projections h A^T of every block (fp32 [n, adapted slabs * rank]), kept for the adapter gradients
        self.hA = [torch.empty(n, lora_cols, device=dev, dtype=f32) for _ in range(depth if save else 1)] if lora_cols else None
        # ln_1 outputs of every block (BF16), kept only when the adapters' down-projections train (SD-LoRA: d lora_A = (dY B)^T h)
        self.hs = [torch.empty(n, DIM, device=dev, dtype=bf) for _ in range(depth)] if (keep_h and save) else None
        self.bwd = None

    def idx(self, layer: int) -> int:
        return layer if self.save else 0

    def prefix_grads(self, layer: int, P: int):
        """fp32 [B, P, 768] gradients of the prefix keys / values of `layer` (written by the attention backward)."""
        d = self.__dict__.setdefault("_dprefix", {})
        if (layer, P) not in d:
            d[(layer, P)] = (torch.empty(self.B, P, DIM, device=self.y.device), torch.empty(self.B, P, DIM, device=self.y.device))
        return d[(layer, P)]

    def backward_buffers(self):
        """Transient buffers of the input-gradient pass (allocated on first use)."""
        if self.bwd is None:
            B, T, Tp, dev = self.B, self.T, self.Tp, self.y.device
            bf, f32 = torch.bfloat16, torch.float32
            n = B * T
            e = lambda *shape, dt=bf: torch.empty(*shape, device=dev, dtype=dt)
            self.bwd = dict(g=[e(B, T, DIM, dt=f32), e(B, T, DIM, dt=f32)], gbf=e(n, DIM), dU=e(n, MLP), dh=e(n, DIM, dt=f32), dO=e(n, DIM),
                            rowdot=e(B, HEADS, T, dt=f32), dqkv=e(n, 3 * DIM), dqkv2=e(n, 3 * DIM))
        return self.bwd


class LoraState:
    """Adapters on `slabs` (0 = q, 1 = k, 2 = v) of the fused QKV projection of every block: A [L][ns][r][D] (fixed during a task), B / dB
    [L][ns][D][r] (views of the owner's flat trainable / gradient arenas)."""

    def __init__(self, engine: "ViTEngine", slabs, rank: int, B: torch.Tensor, dB: torch.Tensor):
        L = engine.depth
        self.slabs, self.ns, self.rank = tuple(slabs), len(slabs), rank
        assert self.ns in (1, 2) and all(0 <= s <= 2 for s in slabs) and list(slabs) == sorted(slabs)
        self.slab_mask = sum(1 << s for s in slabs)
        self.nad, self.R = 1, rank                    # stacked adapters per slab and their total rank
        self.cols = self.ns * rank
        self.ldz = (self.cols + 3) // 4 * 4           # row stride of the saved down-projections (fp32 rows 16-byte aligned)
        self.scale = None
        self.train_A = False
        self.A = torch.zeros(L, self.ns, rank, DIM, device=engine.dev)
        self.A_bf = torch.zeros(L, self.cols, DIM, device=engine.dev, dtype=torch.bfloat16)
        assert B.shape == (L, self.ns, DIM, rank) and dB.shape == B.shape and B.is_contiguous() and dB.is_contiguous()
        self.B, self.dB = B, dB
        self.nchunk = 148
        self.partial = torch.empty(engine.lib.lc_lora_bgrad_partial_floats(self.ns, DIM, rank, self.nchunk), device=engine.dev)
        self.active = False
        self.engine = engine

    def set_A(self, A: torch.Tensor):
        """Installs the down-projections for the coming task (and their BF16 GEMM copy)."""
        self.A.copy_(A.reshape(self.A.shape))
        self.engine._cast(self.A.reshape(self.A_bf.shape), self.A_bf)

    def sync(self):
        """Hook run before every merge (adapters whose parameters change per step refresh their derived copies here)."""


class SDLoraState(LoraState):
    """MultiHeadAttention_SDLoRA (transformer.py:276-357): adapters on q and v, one per task stacked along the rank axis (R = rank * (t + 1)); the
    current one (A and B trainable, magnitude mag[t]) next to the frozen earlier ones, each scaled by (mag[i] + assimilated[i]) / (|B_i| |A_i|); all
    magnitudes train and are shared by the 12 blocks (sd_lora.py:122-125).  A_cur [L][2][r][D], B_cur [L][2][D][r], mag [t + 1] and their gradients are
    views of the owner's flat arenas; A_old / B_old hold the earlier adapters [L][2][t*r][D] / [L][2][D][t*r]."""

    def __init__(self, engine: "ViTEngine", rank: int, nad: int, A_cur, dA_cur, B_cur, dB_cur, mag, dmag, A_old=None, B_old=None, assimilated=None):
        L, dev = engine.depth, engine.dev
        self.engine = engine
        self.slabs, self.ns, self.rank = (0, 2), 2, rank
        self.slab_mask = 0b101
        self.nad, self.R = nad, rank * nad
        assert self.R <= 128, "at most 128 stacked adapter ranks"
        self.Rp = (self.R + 3) // 4 * 4
        self.cols = 2 * self.R
        self.ldz = (self.cols + 3) // 4 * 4
        self.ldg = 2 * self.Rp
        self.train_A = True
        self.A_cur, self.dA_cur, self.B_cur, self.dB_cur, self.mag, self.dmag = A_cur, dA_cur, B_cur, dB_cur, mag, dmag
        assert A_cur.shape == (L, 2, rank, DIM) and B_cur.shape == (L, 2, DIM, rank) and mag.shape == (nad,)
        self.A = torch.zeros(L, 2, self.R, DIM, device=dev)
        self.B = torch.zeros(L, 2, DIM, self.R, device=dev)
        self.A_bf = torch.zeros(L, self.cols, DIM, device=dev, dtype=torch.bfloat16)
        self.BT_bf = torch.zeros(L, 2, self.R, DIM, device=dev, dtype=torch.bfloat16)
        self.scale = torch.ones(L, 2, self.R, device=dev)
        self.inv_norm = torch.ones(L, 2, self.R, device=dev)            # 1 / (|B_i| |A_i|) per column of an old adapter (0 if either norm is 0), 1 for the current
        self.assim = torch.zeros(nad, device=dev) if assimilated is None else assimilated.to(dev)
        r0 = rank * (nad - 1)
        if nad > 1:
            self.A[:, :, :r0].copy_(A_old); self.B[:, :, :, :r0].copy_(B_old)
            for i in range(nad - 1):
                sl = slice(i * rank, (i + 1) * rank)
                nb = self.B[:, :, :, sl].flatten(2).norm(dim=2); na = self.A[:, :, sl].flatten(2).norm(dim=2)         # Frobenius norms, per block and slab
                ok = (nb != 0) & (na != 0)
                self.inv_norm[:, :, sl] = torch.where(ok, 1.0 / (nb * na).clamp_min(1e-30), torch.zeros_like(nb)).unsqueeze(-1)
            self.A_bf.view(L, 2, self.R, DIM)[:, :, :r0].copy_(self.A[:, :, :r0])
            self.BT_bf[:, :, :r0].copy_(self.B[:, :, :, :r0].transpose(2, 3))
        self.cur = slice(r0, self.R)
        self.col_adapter = torch.arange(self.R, device=dev) // rank
        self.nchunk = 148
        self.partial = torch.empty(engine.lib.lc_lora_bgrad_partial_floats(2, DIM, rank, self.nchunk), device=dev)
        self.cd_chunks = 296
        self.cd_partial = torch.empty(self.cd_chunks * self.R, device=dev)
        self._g = {}
        self.active = False

    def gbuf(self, n: int) -> torch.Tensor:
        if n not in self._g:
            self._g[n] = torch.zeros(n, self.ldg, device=self.engine.dev)
        return self._g[n]

    def sync(self):
        """Current adapter -> stacked fp32 / BF16 copies; column scales from the (trainable) magnitudes; zero the magnitude gradient accumulator."""
        L = self.engine.depth
        self.A[:, :, self.cur].copy_(self.A_cur)
        self.B[:, :, :, self.cur].copy_(self.B_cur)
        self.A_bf.view(L, 2, self.R, DIM)[:, :, self.cur].copy_(self.A_cur)
        self.BT_bf[:, :, self.cur].copy_(self.B_cur.transpose(2, 3))
        torch.mul((self.mag + self.assim)[self.col_adapter], self.inv_norm, out=self.scale)
        self.dmag.zero_()


class StackedLoraState(LoraState):
    """`Attention_LoRA` of vit_inflora.py:176-252 (InfLoRA, original): one rank-r adapter per task on k and v, the forward adds the SUM of the adapters
    of tasks 0..t (`weight_k = sum_t B_t A_t`, :236-240); only the current task's lora_B trains.  The merge sees all of them stacked along the rank
    axis (R = r (t + 1)), the gradient path only the current one (exactly InfLoRA_OPT's)."""

    def __init__(self, engine: "ViTEngine", rank: int, nad: int, B_cur: torch.Tensor, dB_cur: torch.Tensor, A_old=None, B_old=None):
        super().__init__(engine, (1, 2), rank, B_cur, dB_cur)
        L, dev = engine.depth, engine.dev
        self.nad, self.R = nad, rank * nad
        assert self.R <= 128, "at most 128 stacked adapter ranks"
        self.B_cur = B_cur
        self.A_stack = torch.zeros(L, 2, self.R, DIM, device=dev)
        self.B_stack = torch.zeros(L, 2, DIM, self.R, device=dev)
        if nad > 1:
            self.A_stack[:, :, :rank * (nad - 1)].copy_(A_old); self.B_stack[:, :, :, :rank * (nad - 1)].copy_(B_old)
        self.cur = slice(rank * (nad - 1), self.R)
        self.A_small = self.A                                   # [L, 2, r, D]: the current adapter (gradient path)
        self.A, self.B = self.A_stack, self.B_stack             # what lc_lora_merge reads

    def set_A(self, A: torch.Tensor):
        self.A_small.copy_(A.reshape(self.A_small.shape))
        self.A_stack[:, :, self.cur].copy_(self.A_small)
        self.engine._cast(self.A_small.reshape(self.A_bf.shape), self.A_bf)

    def sync(self):
        self.B_stack[:, :, :, self.cur].copy_(self.B_cur)


class ViTEngine:
    def __init__(self, depth: int = 12, device=None):
        self.lib = _lib.load()
        self.dev = torch.device(device if device is not None else "cuda:0")
        if self.dev.type != "cuda":
            raise _lib.LcError("ViTEngine needs a CUDA device: there is no CPU path")
        self.depth = depth
        self.layout = vit_param_layout(depth)
        self.w: Dict[str, torch.Tensor] = {}          # fp32 master (frozen)
        self.wb: Dict[str, torch.Tensor] = {}         # bf16 [N, K]   (forward B operand)
        self.wbt: Dict[str, torch.Tensor] = {}        # bf16 [K, N]   (backward-data B operand)
        self.err = torch.zeros(1, device=self.dev, dtype=torch.int32)
        self._ws: Dict[tuple, _Workspace] = {}
        self.launches = 0
        self.lora: Optional[LoraState] = None
        # LayerNorm eps inside the blocks: 1e-5 in transformer.py's ResidualAttentionBlock (:1289,1315), 1e-6 in vit_inflora.py's timm-style Block; the
        # final norm is 1e-6 in both
        self.block_ln_eps = 1e-5
        self.cov: Optional[torch.Tensor] = None       # [L, 768, 768] running sums of h^T h while an input-matrix pass is on (see input_matrix_begin)
        self.cov_rows = 0

    # ---- weights -------------------------------------------------------------------------------
    def load_state(self, state: Dict[str, torch.Tensor]):
        """`state`: the reference `VisionTransformer.state_dict()` (keys as in `vit_param_layout`; `ViTZoo` maps timm names onto them:
        vit.py:70-84)."""
        # the fused QKV weights of all layers share one arena per representation, so that the adapter merge (lc_lora_merge) is one launch
        L = self.depth
        self.qkv_w = torch.empty(L, 3 * DIM, DIM, device=self.dev, dtype=torch.float32)
        self.qkv_wb = torch.empty(L, 3 * DIM, DIM, device=self.dev, dtype=torch.bfloat16)
        self.qkv_wbt = torch.empty(L, DIM, 3 * DIM, device=self.dev, dtype=torch.bfloat16)
        for name, shape in self.layout:
            t = state[name].detach().to(self.dev, torch.float32).contiguous()
            assert tuple(t.shape) == tuple(shape), (name, t.shape, shape)
            if name.endswith("attn.qkv.weight"):
                i = int(name.split(".")[2])
                self.qkv_w[i].copy_(t)
                t = self.qkv_w[i]
            self.w[name] = t
        gemm_w = ["patch_embed.proj.weight"] + [f"transformer.blocks.{i}.{n}.weight" for i in range(self.depth)
                                                 for n in ("attn.qkv", "attn.proj", "mlp.fc1", "mlp.fc2")]
        for name in gemm_w:
            w2 = self.w[name].reshape(self.w[name].shape[0], -1)
            if name.endswith("attn.qkv.weight"):
                i = int(name.split(".")[2])
                self.wb[name], self.wbt[name] = self.qkv_wb[i], self.qkv_wbt[i]
                self._cast(w2, self.wb[name]); self._cast(w2.t().contiguous(), self.wbt[name])
                continue
            self.wb[name] = self._cast(w2)
            if name != "patch_embed.proj.weight":
                self.wbt[name] = self._cast(w2.t().contiguous())
        self.pos_rest = self.w["pos_embed"][0, 1:].contiguous()
        self.pos_cls = self.w["pos_embed"][0, 0].contiguous()
        self.cls = self.w["cls_token"].reshape(DIM).contiguous()

    def _cast(self, t: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        if out is None:
            out = torch.empty(t.shape, device=self.dev, dtype=torch.bfloat16)
        assert t.is_contiguous() and out.is_contiguous() and out.shape == t.shape
        check(self.lib.lc_cast_bf16(t.data_ptr(), out.data_ptr(), t.numel(), stream_ptr()), "cast_bf16")
        return out

    def workspace(self, B: int, T: int, save: bool) -> _Workspace:
        lc = self.lora.ldz if self.lora is not None else 0
        keep_h = self.lora is not None and self.lora.train_A
        key = (B, T, save, lc, keep_h)
        if key not in self._ws:
            self._ws[key] = _Workspace(B, T, self.depth, save, self.dev, lora_cols=lc, keep_h=keep_h)
        return self._ws[key]

    def tensor_core_error(self) -> bool:
        """True if any tcgen05 kernel hit an mbarrier timeout (it then wrote nothing useful)."""
        return bool(self.err.item())

    # ---- GEMM plumbing -------------------------------------------------------------------------
    def gemm(self, A, lda, B, ldb, C, ldc, M, N, K, *, sA=(0, 0), sB=(0, 0), sC=(0, 0), batch=(1, 1), bias=None, residual=None, ldr=0, sR=(0, 0),
             out2=None, gelu_bwd_aux=None, out_f32=False, alpha=1.0, gelu_mode=0, ksplit=0, stride_split=0):
        """C[z] = alpha * A[z] @ B[z]^T (+bias +residual); A/B/C/residual/out2 are raw device addresses (ints), strides in elements.
        ksplit > 1: the K blocks are dealt to `ksplit` fp32 partial outputs, partial s at C + s * stride_split elements (no epilogue operands)."""
        d = GemmDesc()
        d.A, d.lda, d.strideA_in, d.strideA_out = A, lda, sA[0], sA[1]
        d.B, d.ldb, d.strideB_in, d.strideB_out = B, ldb, sB[0], sB[1]
        d.C, d.ldc, d.strideC_in, d.strideC_out = C, ldc, sC[0], sC[1]
        d.bias, d.residual, d.ldr, d.strideR_in, d.strideR_out = bias, residual, ldr, sR[0], sR[1]
        d.out2 = out2
        d.gelu_bwd_aux = gelu_bwd_aux
        d.M, d.N, d.K, d.batch_in, d.batch_out, d.out_f32, d.alpha = M, N, K, batch[0], batch[1], int(out_f32), alpha
        d.gelu_mode = gelu_mode
        d.ksplit, d.strideC_split = ksplit, stride_split
        check(self.lib.lc_gemm_bf16_ex(ctypes.byref(d), self.err.data_ptr(), stream_ptr()), f"gemm {M}x{N}x{K}")
        self.launches += 1

    def _linear(self, a_bf16: torch.Tensor, wname: str, out: torch.Tensor, *, bias=None, residual=None, out2=None, rows=None, gelu_mode=0):
        w = self.wb[wname]
        N, K = w.shape
        M = a_bf16.shape[0] if rows is None else rows
        self.gemm(a_bf16.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, M, N, K, bias=None if bias is None else self.w[bias].data_ptr(),
                  residual=None if residual is None else residual.data_ptr(), ldr=N, out2=None if out2 is None else out2.data_ptr(),
                  out_f32=out.dtype == torch.float32, gelu_mode=gelu_mode)

    def _ln(self, x: torch.Tensor, wname: str, eps: float, out_bf16=None, out_f32=None, stat=None):
        rows = x.numel() // DIM
        check(self.lib.lc_layernorm_forward(x.data_ptr(), self.w[wname + ".weight"].data_ptr(), self.w[wname + ".bias"].data_ptr(), eps, rows, DIM,
                                            None if out_bf16 is None else out_bf16.data_ptr(), None if out_f32 is None else out_f32.data_ptr(),
                                            None if stat is None else stat.data_ptr(), stream_ptr()), "layernorm")
        self.launches += 1

    # ---- forward -------------------------------------------------------------------------------
    def embed(self, img: torch.Tensor, prompts: Optional[torch.Tensor], ws: _Workspace):
        """Tokens [prompts (shared by the batch) | cls + pos[0] | patches + pos[1:]] into ws.x[0]."""
        B, T = ws.B, ws.T
        Pn = 0 if prompts is None else prompts.shape[0]
        assert T == Pn + PATCHES + 1 and img.shape == (B, 3, 224, 224) and img.is_contiguous() and img.dtype == torch.float32
        st = stream_ptr()
        x0 = ws.x[0]
        check(self.lib.lc_vit_patchify(img.data_ptr(), ws.patches.data_ptr(), B, st), "patchify")
        w = self.wb["patch_embed.proj.weight"]
        self.gemm(ws.patches.data_ptr(), DIM, w.data_ptr(), DIM, x0.data_ptr() + (Pn + 1) * DIM * 4, DIM, PATCHES, DIM, DIM, sA=(PATCHES * DIM, 0),
                  sC=(T * DIM, 0), batch=(B, 1), bias=self.w["patch_embed.proj.bias"].data_ptr(), residual=self.pos_rest.data_ptr(), ldr=DIM, out_f32=True)
        check(self.lib.lc_vit_set_rows(x0.data_ptr(), T * DIM, B, Pn, 1, self.cls.data_ptr(), self.pos_cls.data_ptr(), DIM, st), "cls row")
        if Pn:
            assert prompts.is_contiguous() and prompts.dtype == torch.float32 and prompts.shape[1] == DIM
            check(self.lib.lc_vit_set_rows(x0.data_ptr(), T * DIM, B, 0, Pn, prompts.data_ptr(), None, DIM, st), "prompt rows")
        self.launches += 3 + (1 if Pn else 0)

    def block_forward(self, i: int, ws: _Workspace, prefix=None):
        """`prefix` = (pk, pv) BF16 [B, P, 768]: prefix keys / values of this block (transformer.py:175-180), or None."""
        B, T, Tp = ws.B, ws.T, ws.Tp
        st = stream_ptr()
        pre = f"transformer.blocks.{i}."
        k = ws.idx(i)
        xin = ws.x[i] if ws.save else ws.x[i % 2]
        xout = ws.x[i + 1] if ws.save else ws.x[(i + 1) % 2]
        xmid, qkv, o, upre = ws.xmid[k], ws.qkv[k], ws.o[k], ws.upre[k]
        h1 = ws.hs[i] if ws.hs is not None else ws.h
        self._ln(xin, pre + "ln_1", self.block_ln_eps, out_bf16=h1)
        if self.cov is not None:
            self._accumulate_input_matrix(i, ws)
        if self.lora is not None and self.lora.active and ws.save:
            lo = self.lora
            self.gemm(h1.data_ptr(), DIM, lo.A_bf[i].data_ptr(), DIM, ws.hA[k].data_ptr(), lo.ldz, B * T, lo.cols, DIM, out_f32=True)
        self._linear(h1, pre + "attn.qkv.weight", qkv, bias=pre + "attn.qkv.bias")
        if prefix is None:
            check(self.lib.lc_attn_forward(qkv.data_ptr(), o.data_ptr(), ws.lse[k].data_ptr(), B, T, HEADS, self.err.data_ptr(), st), "attn_forward")
        else:
            pk, pv = prefix
            assert pk.dtype == torch.bfloat16 and pk.shape == pv.shape and pk.shape[0] == B and pk.shape[2] == DIM and pk.is_contiguous() and pv.is_contiguous()
            check(self.lib.lc_attn_forward_prefix(qkv.data_ptr(), o.data_ptr(), ws.lse[k].data_ptr(), B, T, HEADS, pk.data_ptr(), pv.data_ptr(), pk.shape[1],
                                                  self.err.data_ptr(), st), "attn_forward_prefix")
        self._linear(o, pre + "attn.proj.weight", xmid, bias=pre + "attn.proj.bias", residual=xin)
        self._ln(xmid, pre + "ln_2", self.block_ln_eps, out_bf16=ws.h)
        # saved passes keep GELU'(fc1 output) (all the backward needs of it); no-grad passes keep nothing but GELU(fc1 output)
        self._linear(ws.h, pre + "mlp.fc1.weight", upre, bias=pre + "mlp.fc1.bias", out2=ws.u, gelu_mode=1 if ws.save else 2)
        self._linear(ws.u, pre + "mlp.fc2.weight", xout, bias=pre + "mlp.fc2.bias", residual=xmid)
        self.launches += 1

    def forward(self, img: torch.Tensor, prompts: Optional[torch.Tensor] = None, save: bool = False, prefix: Optional[Dict[int, tuple]] = None) -> _Workspace:
        """Runs the backbone; returns the workspace (ws.y = final-LayerNorm tokens fp32 [B, T, 768]).  `prefix[i]` = (pk, pv) of block i."""
        B = img.shape[0]
        T = PATCHES + 1 + (0 if prompts is None else prompts.shape[0])
        ws = self.workspace(B, T, save)
        self.embed(img, prompts, ws)
        ws.prefix = prefix if save else None
        for i in range(self.depth):
            self.block_forward(i, ws, None if prefix is None else prefix.get(i))
        last = ws.x[self.depth] if save else ws.x[self.depth % 2]
        self._ln(last, "norm", 1e-6, out_f32=ws.y, stat=ws.ystat if save else None)
        return ws

    def pooled(self, ws: _Workspace, n_prompt: int) -> torch.Tensor:
        """Mean over the prompt positions of the normalised tokens (transformer.py:2255-2258), or the cls row without prompts (:2260)."""
        check(self.lib.lc_vit_pool_rows(ws.y.data_ptr(), ws.T * DIM, ws.B, 0, max(n_prompt, 1), DIM, ws.feat.data_ptr(), stream_ptr()), "pool_rows")
        self.launches += 1
        return ws.feat

    def linear_head(self, feat: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], logits: torch.Tensor):
        B, C = feat.shape[0], W.shape[0]
        check(self.lib.lc_linear_head(feat.data_ptr(), W.data_ptr(), None if bias is None else bias.data_ptr(), B, C, DIM, logits.data_ptr(),
                                      logits.stride(0), stream_ptr()), "linear_head")
        self.launches += 1
        return logits

    # ---- backward wrt the input tokens -----------------------------------------------------------
    def _linear_t(self, a_bf16: torch.Tensor, wname: str, out: torch.Tensor, gelu_bwd_aux=None):
        """out = a @ W  (backward-data of y = x W^T): the B operand is the pre-transposed BF16 copy [in_features][out_features]."""
        w = self.wbt[wname]
        N, K = w.shape
        self.gemm(a_bf16.data_ptr(), K, w.data_ptr(), K, out.data_ptr(), N, a_bf16.shape[0], N, K, out_f32=out.dtype == torch.float32,
                  gelu_bwd_aux=None if gelu_bwd_aux is None else gelu_bwd_aux.data_ptr(), gelu_mode=0 if gelu_bwd_aux is None else 1)

    def _ln_bwd(self, dh, x, wname, eps, res, out_f32, out_bf16, dh_pool=None, T=0, n_active=0):
        rows = x.numel() // DIM
        check(self.lib.lc_layernorm_backward(None if dh is None else dh.data_ptr(), None if dh_pool is None else dh_pool.data_ptr(), T, n_active, x.data_ptr(),
                                             self.w[wname + ".weight"].data_ptr(), eps, rows, DIM, None if res is None else res.data_ptr(),
                                             None if out_f32 is None else out_f32.data_ptr(), None if out_bf16 is None else out_bf16.data_ptr(), stream_ptr()),
              "layernorm_backward")
        self.launches += 1

    def backward_tokens(self, ws: _Workspace, dfeat: torch.Tensor, n_prompt: int, to_tokens: bool = True) -> Optional[torch.Tensor]:
        """Given d(loss)/d(pooled feature) [B, 768], returns d(loss)/d(input tokens) fp32 [B, T, 768] (ws must come from forward(save=True)).
        Autograd of transformer.py:2006-2017 / :1331-1336 / :169-197 restricted to the activations: no weight gradients (frozen backbone)."""
        assert ws.save, "backward needs the activations of forward(save=True)"
        B, T, Tp = ws.B, ws.T, ws.Tp
        st = stream_ptr()
        bw = ws.backward_buffers()
        g, g2 = bw["g"]
        gbf, dU, dh, dO, rowdot = (bw[k] for k in ("gbf", "dU", "dh", "dO", "rowdot"))
        # Adapter gradients are leaves of the dependency graph: they run on a second stream next to the token-gradient chain (the persistent GEMMs leave
        # registers / issue slots for one CUDA-core CTA per SM).  d(qkv) is double-buffered so that block i-1's attention backward may write one buffer
        # while block i's adapter kernels still read the other; events carry the edges, and fork / join keep the pair capturable into one CUDA graph.
        side_on = self.lora is not None and self.lora.active
        dq = [bw["dqkv"], bw["dqkv2"]] if side_on else [bw["dqkv"], bw["dqkv"]]
        if side_on:
            main = torch.cuda.current_stream(self.dev)
            if getattr(self, "_side", None) is None:
                self._side = torch.cuda.Stream(device=self.dev)
                self._ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
                self._ev_done = [torch.cuda.Event(), torch.cuda.Event()]
            used = [False, False]
            self._side.wait_stream(main)
        self._ln_bwd(None, ws.x[self.depth], "norm", 1e-6, None, g, gbf, dh_pool=dfeat, T=T, n_active=max(n_prompt, 1))
        for i in reversed(range(self.depth)):
            pre = f"transformer.blocks.{i}."
            kb = i & 1
            dqkv = dq[kb]
            if side_on and used[kb]:
                main.wait_event(self._ev_done[kb])          # the adapter kernels that last read this buffer
            # MLP branch: x_out = x_mid + fc2(GELU(fc1(LN2(x_mid))))
            self._linear_t(gbf, pre + "mlp.fc2.weight", dU, gelu_bwd_aux=ws.upre[i])
            self._linear_t(dU, pre + "mlp.fc1.weight", dh)
            self._ln_bwd(dh, ws.xmid[i], pre + "ln_2", self.block_ln_eps, g, g2, gbf)
            # attention branch: x_mid = x_in + proj(softmax(QK^T/8) V)
            self._linear_t(gbf, pre + "attn.proj.weight", dO)
            pfx = None if getattr(ws, "prefix", None) is None else ws.prefix.get(i)
            if pfx is None:
                check(self.lib.lc_attn_backward(ws.qkv[i].data_ptr(), ws.o[i].data_ptr(), dO.data_ptr(), ws.lse[i].data_ptr(), rowdot.data_ptr(), dqkv.data_ptr(),
                                                B, T, HEADS, self.err.data_ptr(), st), "attn_backward")
            else:
                pk, pv = pfx
                dpk, dpv = ws.prefix_grads(i, pk.shape[1])
                check(self.lib.lc_attn_backward_prefix(ws.qkv[i].data_ptr(), dO.data_ptr(), ws.lse[i].data_ptr(), dqkv.data_ptr(), B, T, HEADS, pk.data_ptr(),
                                                       pv.data_ptr(), dpk.data_ptr(), dpv.data_ptr(), pk.shape[1], self.err.data_ptr(), st), "attn_backward_prefix")
            self.launches += 2
            if side_on:
                self._ev_ready[kb].record(main)
                with torch.cuda.stream(self._side):
                    self._side.wait_event(self._ev_ready[kb])
                    self.lora_grads(i, ws, dqkv)
                    self._ev_done[kb].record(self._side)
                used[kb] = True
            if i == 0 and not to_tokens:          # nothing trainable below the first attention (adapters only): stop here
                if side_on:
                    main.wait_stream(self._side)
                return None
            self._linear_t(dqkv, pre + "attn.qkv.weight", dh)
            self._ln_bwd(dh, ws.x[i], pre + "ln_1", self.block_ln_eps, g2, g, gbf)
        if side_on:
            main.wait_stream(self._side)
        return g

    # ---- InfLoRA's input matrix (transformer.py:242-244, InfLoRA_opt.py:243-245) ----------------------
    def input_matrix_begin(self):
        self.cov = torch.zeros(self.depth, DIM, DIM, device=self.dev)
        self.cov_rows = 0

    def input_matrix_end(self) -> torch.Tensor:
        """Mean over all tokens seen since input_matrix_begin() of h h^T, h = ln_1(x): [L, 768, 768] (the reference's `cur_matrix` per block)."""
        out = self.cov / float(max(self.cov_rows, 1))
        self.cov = None
        return out

    def _accumulate_input_matrix(self, i: int, ws: _Workspace):
        n = ws.B * ws.T
        npad = _up8(n)
        if getattr(ws, "hT", None) is None:
            ws.hT = torch.empty(DIM, npad, device=self.dev, dtype=torch.bfloat16)
        check(self.lib.lc_transpose_bf16(ws.h.data_ptr(), DIM, n, DIM, ws.hT.data_ptr(), npad, stream_ptr()), "transpose_bf16")
        # K = all tokens of the batch against a 768 x 768 output: 18 output tiles, so the K blocks are dealt to COV_SPLIT partial outputs (144 CTAs on 148
        # SMs instead of 18) and summed in a fixed order
        if getattr(ws, "cov_part", None) is None:
            ws.cov_part = torch.empty(COV_SPLIT, DIM, DIM, device=self.dev)
        nkb, ks = (npad + 63) // 64, COV_SPLIT
        while ks > 1 and (ks - 1) * ((nkb + ks - 1) // ks) >= nkb:      # every split must own at least one K block (lc_gemm_bf16_ex checks the same)
            ks -= 1
        self.gemm(ws.hT.data_ptr(), npad, ws.hT.data_ptr(), npad, ws.cov_part.data_ptr(), DIM, DIM, DIM, npad, out_f32=True, ksplit=ks, stride_split=DIM * DIM)
        self.cov[i] += ws.cov_part[:ks].sum(0)
        self.launches += 1
        if i == 0:
            self.cov_rows += n

    # ---- low-rank adapters on the QKV projection ----------------------------------------------------
    def lora_merge(self, w_out: bool = False):
        """W' = W + B A for the adapted slabs of every layer into the BF16 GEMM operands (w_out: also into the fp32 master = `merge_weight`)."""
        lo = self.lora
        lo.sync()
        check(self.lib.lc_lora_merge(self.qkv_w.data_ptr(), lo.A.data_ptr(), lo.B.data_ptr(), None if lo.scale is None else lo.scale.data_ptr(), lo.slab_mask,
                                     self.depth, DIM, lo.R,
                                     self.qkv_wb.data_ptr(), self.qkv_wbt.data_ptr(), self.qkv_w.data_ptr() if w_out else None, stream_ptr()), "lora_merge")
        self.launches += 1

    def lora_grads(self, i: int, ws: _Workspace, dqkv: torch.Tensor):
        """d lora_B[i] = d(slab)^T (h A^T): rank-form adapter gradient of block i from the token gradient of the adapted QKV slabs."""
        lo = self.lora
        n = ws.B * ws.T
        st = stream_ptr()
        xs = (lo.slabs[1] - lo.slabs[0]) * DIM if lo.ns > 1 else DIM
        if not lo.train_A:
            check(self.lib.lc_lora_bgrad_rows(dqkv.data_ptr(), 3 * DIM, lo.slabs[0] * DIM, xs, lo.ns, DIM, ws.hA[i].data_ptr(), lo.ldz, lo.rank, n, lo.partial.data_ptr(),
                                              lo.nchunk, lo.dB[i].data_ptr(), st), "lora_bgrad_rows")
            self.launches += 2
            return
        # SD-LoRA: G = d(slab) B_cat ([n, R] per slab) -> magnitudes; current adapter: dB = mag dY^T (h A^T), dA = mag (dY B)^T h
        G = lo.gbuf(n)
        z0 = (lo.nad - 1) * lo.rank
        for si, slab in enumerate(lo.slabs):
            self.gemm(dqkv.data_ptr() + 2 * slab * DIM, 3 * DIM, lo.BT_bf[i, si].data_ptr(), DIM, G.data_ptr() + 4 * si * lo.Rp, lo.ldg, n, lo.R, DIM, out_f32=True)
            check(self.lib.lc_coldot_accumulate(G.data_ptr() + 4 * si * lo.Rp, lo.ldg, ws.hA[i].data_ptr(), lo.ldz, si * lo.R, lo.R, lo.rank, n,
                                                lo.inv_norm[i, si].data_ptr(), lo.cd_partial.data_ptr(), lo.cd_chunks, lo.dmag.data_ptr(), st), "coldot")
        mag_cur = lo.mag.data_ptr() + 4 * (lo.nad - 1)
        check(self.lib.lc_rowouter_bf16(dqkv.data_ptr(), 3 * DIM, lo.slabs[0] * DIM, xs, lo.ns, DIM, ws.hA[i].data_ptr(), lo.ldz, z0, lo.R, lo.rank, n,
                                        lo.partial.data_ptr(), lo.nchunk, lo.dB_cur[i].data_ptr(), 0, mag_cur, st), "rowouter dB")
        check(self.lib.lc_rowouter_bf16(ws.hs[i].data_ptr(), DIM, 0, 0, lo.ns, DIM, G.data_ptr(), lo.ldg, z0, lo.Rp, lo.rank, n,
                                        lo.partial.data_ptr(), lo.nchunk, lo.dA_cur[i].data_ptr(), 1, mag_cur, st), "rowouter dA")
        self.launches += 4 * lo.ns + 4

    def prompt_row_grads(self, g: torch.Tensor, n_prompt: int, out: torch.Tensor) -> torch.Tensor:
        """Gradient of the prompt rows shared by the batch: out[r] = sum_b g[b, r]  (the `repeat(B, 1)` of prompt.py:390)."""
        B, T = g.shape[0], g.shape[1]
        check(self.lib.lc_sum_batch_rows(g.data_ptr(), T * DIM, B, n_prompt, DIM, out.data_ptr(), stream_ptr()), "sum_batch_rows")
        self.launches += 1
        return out
