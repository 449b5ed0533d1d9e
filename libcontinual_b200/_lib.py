"""ctypes binding of the C-ABI CUDA library (`include/lc_b200.h`).  No fallback: if the shared object is missing or a
call fails the product path raises."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int64, c_longlong, c_uint32, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "liblc_b200.so")

_lib = None

# name -> (restype, argtypes).  Kept in one table so the symbol-export test can walk it.
P = c_void_p
SIGNATURES = {
    "lc_version": (c_char_p, []),
    "lc_device_check": (c_int, []),
    "lc_resnet_create": (c_int, [c_int, c_int, c_int, c_int, POINTER(c_void_p)]),
    "lc_resnet_destroy": (None, [P]),
    "lc_resnet_set_mode": (c_int, [P, c_int]),
    "lc_resnet_get_mode": (c_int, [P]),
    "lc_resnet_set_last_relu": (c_int, [P, c_int]),
    "lc_resnet_param_count": (c_longlong, [P]),
    "lc_resnet_rstat_count": (c_longlong, [P]),
    "lc_resnet_workspace_floats": (c_longlong, [P]),
    "lc_resnet_num_convs": (c_int, [P]),
    "lc_resnet_num_launches": (c_int, [P, c_int]),
    "lc_resnet_conv_info": (c_int, [P, c_int, POINTER(c_longlong), POINTER(c_int), POINTER(c_int), POINTER(c_int), POINTER(c_int),
                                    POINTER(c_longlong), POINTER(c_longlong), POINTER(c_longlong)]),
    "lc_resnet_ws_offset": (c_longlong, [P, c_int]),
    "lc_resnet_forward": (c_int, [P, P, c_int, P, P, P, c_int, c_int, P]),
    "lc_resnet_backward": (c_int, [P, P, c_int, P, P, P, P]),
    "lc_head_forward": (c_int, [P, c_int, c_int, c_int, P, P, c_int, P, P, c_int, P]),
    "lc_loss_ce_kd": (c_int, [P, c_int, P, c_int, P, c_int, c_int, c_int, c_int, c_float, c_float, c_int, P, P, P, P]),
    "lc_head_backward": (c_int, [P, c_int, P, P, c_int, c_int, c_int, P, P, P, P, c_int, P]),
    "lc_avgpool_forward": (c_int, [P, c_int, c_int, c_int, P, P]),
    "lc_avgpool_backward": (c_int, [P, c_int, c_int, c_int, P, P]),
    "lc_ewc_penalty_grad": (c_int, [P, P, P, P, c_longlong, P, P, P, P, P]),
    "lc_fisher_accumulate": (c_int, [P, P, c_longlong, c_float, P]),
    "lc_fisher_merge": (c_int, [P, P, c_longlong, c_float, c_float, P]),
    "lc_sgd_momentum": (c_int, [P, P, P, c_longlong, P, P]),
    "lc_sgd_momentum_frozen": (c_int, [P, P, P, c_longlong, P, c_longlong, c_longlong, P]),
    "lc_adam": (c_int, [P, P, P, P, c_longlong, P, P]),
    "lc_adam_tick": (c_int, [P, P]),
    "lc_clip_grad_norm": (c_int, [P, c_longlong, c_float, P, P, P]),
    "lc_cosine_head_forward": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, c_int, P]),
    "lc_cosine_head_backward": (c_int, [P, c_int, P, P, P, c_int, c_int, c_int, P, P, P]),
    "lc_lucir_loss": (c_int, [P, P, c_int, P, P, c_int, P, c_int, c_int, c_int, c_int, c_float, c_float, c_float, P, P, P, P, P, P, P]),
    "lc_l2p_select": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P]),
    "lc_l2p_select_phase": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_int, P]),
    "lc_l2p_gather": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P]),
    "lc_gpm_project": (c_int, [P, P, c_int, c_int, P]),
    "lc_lora_merge_qkv": (c_int, [P, P, P, P, P, P, c_int, c_int, P]),
    "lc_lora_bgrad": (c_int, [P, P, P, c_int, c_int, P]),
    "lc_prompt_key_match": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "lc_coda_prompt_forward": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "lc_coda_prompt_backward": (c_int, [P, P, P, P, P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "lc_gather_rows_bf16": (c_int, [P, P, c_longlong, c_int, c_int, c_int, P, P]),
    "lc_split_bf16": (c_int, [P, P, P, c_longlong, P]),
    "lc_gpm_project_tc": (c_int, [P, P, P, c_int, c_int, P, P, P, P]),
    "lc_transpose_bf16": (c_int, [P, c_longlong, c_longlong, c_int, P, c_longlong, P]),
    "lc_rowouter_bf16": (c_int, [P, c_longlong, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, c_longlong, P, c_int, P, c_int, P, P]),
    "lc_coldot_accumulate": (c_int, [P, c_int, P, c_int, c_int, c_int, c_int, c_longlong, P, P, c_int, P, P]),
    "lc_lora_merge": (c_int, [P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "lc_lora_bgrad_partial_floats": (c_longlong, [c_int, c_int, c_int, c_int]),
    "lc_lora_bgrad_rows": (c_int, [P, c_longlong, c_int, c_int, c_int, c_int, P, c_int, c_int, c_longlong, P, c_int, P, P]),
    "lc_herding_select": (c_int, [P, P, c_int, c_int, c_int, P, P, P]),
    "lc_ncm_classify": (c_int, [P, P, c_int, c_int, c_int, P, P]),
    "lc_gemm_bf16": (c_int, [P, c_int, c_longlong, P, c_int, c_longlong, P, c_int, c_longlong, c_int, c_int, c_int, c_int, P, P, c_int, c_longlong, P, c_int,
                             c_float, P, P]),
    "lc_gemm_bf16_ex": (c_int, [P, P, P]),
    "lc_conv_gemm_bf16": (c_int, [P, P, P]),
    "lc_nn_bn_scratch_floats": (c_longlong, [c_int]),
    "lc_nn_im2col": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_longlong, P, c_longlong, c_int, P]),
    "lc_nn_im2col_f32": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_longlong, P, c_longlong, P]),
    "lc_nn_col2im": (c_int, [P, c_longlong, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "lc_nn_bn_stats": (c_int, [P, c_longlong, c_int, P, P, c_float, c_float, P, P, P, P]),
    "lc_nn_bn_eval_affine": (c_int, [P, c_int, P, P, c_float, P, P]),
    "lc_nn_bn_act": (c_int, [P, P, P, P, c_longlong, c_int, c_int, c_float, P, c_int, P, P, P]),
    "lc_nn_dropout_mask": (c_int, [P, c_int, c_float, c_longlong, P, P]),
    "lc_nn_rng_advance": (c_int, [P, P]),
    "lc_nn_bn_backward": (c_int, [P, P, P, c_float, P, P, c_longlong, c_int, P, P, P, P, P, P, P]),
    "lc_nn_maxpool_forward": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "lc_nn_maxpool_backward": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "lc_nn_avgpool_forward": (c_int, [P, c_int, c_int, c_int, P, P]),
    "lc_nn_avgpool_backward": (c_int, [P, c_int, c_int, c_int, P, P]),
    "lc_nn_pack_weight": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, c_longlong, P]),
    "lc_nn_wgrad_reduce": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_longlong, P, P]),
    "lc_nn_cast_transpose": (c_int, [P, c_longlong, c_int, P, c_longlong, P, c_longlong, P]),
    "lc_attn_forward": (c_int, [P, P, P, c_int, c_int, c_int, P, P]),
    "lc_attn_backward": (c_int, [P, P, P, P, P, P, c_int, c_int, c_int, P, P]),
    "lc_attn_forward_prefix": (c_int, [P, P, P, c_int, c_int, c_int, P, P, c_int, P, P]),
    "lc_attn_backward_prefix": (c_int, [P, P, P, P, c_int, c_int, c_int, P, P, P, P, c_int, P, P]),
    "lc_vit_patchify": (c_int, [P, P, c_int, P]),
    "lc_vit_set_rows": (c_int, [P, c_longlong, c_int, c_int, c_int, P, P, c_int, P]),
    "lc_layernorm_forward": (c_int, [P, P, P, c_float, c_longlong, c_int, P, P, P, P]),
    "lc_layernorm_backward": (c_int, [P, P, c_int, c_int, P, P, c_float, c_longlong, c_int, P, P, P, P]),
    "lc_sum_batch_rows": (c_int, [P, c_longlong, c_int, c_int, c_int, P, P]),
    "lc_vit_pool_rows": (c_int, [P, c_longlong, c_int, c_int, c_int, c_int, P, P]),
    "lc_linear_head": (c_int, [P, P, P, c_int, c_int, c_int, P, c_int, P]),
    "lc_cast_bf16": (c_int, [P, P, c_longlong, P]),
    "lc_linear_head_backward": (c_int, [P, c_int, P, P, c_int, c_int, c_int, P, P, P, P]),
    "lc_l2p_backward": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, c_float, P, P]),
    "lc_loss_ce_masked": (c_int, [P, c_int, P, c_int, c_int, c_int, P, c_float, P, P, P, P]),
    "lc_conv_scratch_floats": (c_longlong, [c_int, c_int, c_int, c_int]),
    "lc_conv3x3": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P]),
    "lc_conv3x3_packed": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "lc_conv_tc_scratch_floats": (c_longlong, [c_int, c_int, c_int]),
    "lc_conv3x3_tc": (c_int, [P, P, P, c_int, c_int, c_int, c_int, P, P, P, P, P, P, P, P, P]),
    "lc_conv3x3_tc_packed": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, P]),
    "lc_conv3x3s2_tc": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, P, P, P]),
    "lc_conv3x3s2_dgrad_tc": (c_int, [P, P, P, c_int, c_int, c_int, P, P]),
    "lc_conv3x3_wgrad": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, P]),
    "lc_conv3x3_wgrad_tc": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P, P]),
    "lc_conv1x1s2": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P]),
    "lc_bn_act_forward": (c_int, [P, P, P, P, P, P, P, c_longlong, c_int, P]),
    "lc_augment_cifar_u8": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P]),
    "lc_resize_scratch_bytes": (c_longlong, [c_int, c_int, c_int]),
    "lc_resize_crop_u8": (c_int, [P, P, P, P, P, c_int, c_int, c_int, c_int, P, P, P, P]),
    "lc_eval_meter": (c_int, [P, P, c_int, P, c_int, c_int, P, P]),
    "lc_eval_fold": (c_int, [P, P, c_int, c_int, P]),
    "lc_bn_backward": (c_int, [P, P, c_int, P, P, P, P, P, P, c_longlong, c_int, P, P]),
}


class GemmDesc(ctypes.Structure):
    """`lc_gemm_desc` of include/lc_b200.h."""
    _fields_ = [("A", c_void_p), ("lda", c_longlong), ("strideA_in", c_longlong), ("strideA_out", c_longlong),
                ("B", c_void_p), ("ldb", c_longlong), ("strideB_in", c_longlong), ("strideB_out", c_longlong),
                ("C", c_void_p), ("ldc", c_longlong), ("strideC_in", c_longlong), ("strideC_out", c_longlong),
                ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_longlong), ("strideR_in", c_longlong), ("strideR_out", c_longlong),
                ("out2", c_void_p), ("gelu_bwd_aux", c_void_p),
                ("M", c_int), ("N", c_int), ("K", c_int), ("batch_in", c_int), ("batch_out", c_int), ("out_f32", c_int),
                ("alpha", c_float), ("gelu_mode", c_int), ("ksplit", c_int), ("strideC_split", c_longlong)]


class ConvDesc(ctypes.Structure):
    """`lc_conv_desc` of include/lc_b200.h."""
    _fields_ = [("X", c_void_p), ("Wk", c_void_p), ("Y", c_void_p), ("bias", c_void_p), ("residual", c_void_p), ("ldc", c_longlong), ("ldr", c_longlong),
                ("N", c_int), ("H", c_int), ("W", c_int), ("C", c_int), ("Cout", c_int), ("ks", c_int), ("stride", c_int), ("pad", c_int),
                ("Ho", c_int), ("Wo", c_int), ("out_f32", c_int)]


class LcError(RuntimeError):
    pass


def load():
    """Loads the library once.  Raises if it has not been built (`python -m libcontinual_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("LC_B200_LIB", LIB_PATH)      # debug builds (tools/) may point at an instrumented copy
    if not os.path.exists(path):
        raise LcError(f"{path} is missing: build it with `python -m libcontinual_b200.build` (nvcc, sm_100a). "
                      "There is no CPU / PyTorch fallback for the hot path.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        raise LcError(f"libcontinual_b200 call failed ({what}): code {rc}" + (" [invalid argument / unsupported shape]" if rc == -22 else " [CUDA error]"))


def ptr(t):
    """Device pointer of a tensor (or NULL)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device-resident contiguous tensors only"
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def host_acc(model, count, n):
    """Accuracy as the Python float the reference returns from `inference` (a device->host sync, finetune.py:33-36) — unless the caller runs
    `libcontinual_b200.trainer.validate`, which sets `model._defer_metrics` and keeps the counts on the device (lc_eval_meter): then NaN is returned and
    nothing synchronises."""
    if getattr(model, "_defer_metrics", False):
        return float("nan")
    return float(count.item()) / n
