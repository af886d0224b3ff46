"""ctypes binding of libtg_b200.so — the C ABI declared in include/terragan_b200.h.

The library is built in-tree by `__graft_entry__.build()` / `make -C terra-gan_b200/csrc`. There is
deliberately NO fallback: if the shared object is missing or a call fails, a RuntimeError is raised
(the north star forbids a CPU / eager fallback on the hot path).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtg_b200.so")

TG_MAX_TAPS = 64
TG_MAX_SUB = 4
ACT_NONE, ACT_RELU, ACT_LEAKY = 0, 1, 2
DTYPE_BF16, DTYPE_F32 = 0, 1

c_void_p, c_int, c_long, c_float, c_size_t = C.c_void_p, C.c_int, C.c_long, C.c_float, C.c_size_t


class ConvSub(C.Structure):
    _fields_ = [("tap_begin", C.c_int32), ("tap_count", C.c_int32), ("k_off", C.c_int32),
                ("out_plane", C.c_int32)]


class ConvArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("B", C.c_int32), ("P", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("C", C.c_int32),
        ("w", c_void_p), ("N", C.c_int32), ("Ktot", C.c_int32),
        ("num_sub", C.c_int32), ("sub", ConvSub * TG_MAX_SUB),
        ("num_taps", C.c_int32),
        ("tap_plane", C.c_int8 * TG_MAX_TAPS), ("tap_dh", C.c_int8 * TG_MAX_TAPS),
        ("tap_dw", C.c_int8 * TG_MAX_TAPS),
        ("out", c_void_p), ("Po", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32),
        ("code", c_void_p), ("lut", C.POINTER(C.c_float)), ("lut_len", C.c_int32),
        ("bias", c_void_p), ("scale", c_void_p), ("shift", c_void_p),
        ("act", C.c_int32), ("slope", C.c_float),
        ("stats", c_void_p), ("stats_rows_cap", C.c_int32), ("stats_rows_used", C.c_int32),
        ("gate", c_void_p), ("gate_slope", C.c_float),
        ("dtype", C.c_int32), ("addend", c_void_p),
        ("pool_out", c_void_p), ("skip_out", C.c_int32),
    ]


class GradSrc(C.Structure):
    _fields_ = [("ptr", c_void_p), ("pix_stride", C.c_int64), ("chan_off", C.c_int32), ("split", C.c_int32)]


class WgradArgs(C.Structure):
    _fields_ = [
        ("x", c_void_p), ("B", C.c_int32), ("P", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("C", C.c_int32),
        ("g", c_void_p), ("Ho", C.c_int32), ("Wo", C.c_int32), ("N", C.c_int32),
        ("num_taps", C.c_int32),
        ("tap_plane", C.c_int8 * TG_MAX_TAPS), ("tap_dh", C.c_int8 * TG_MAX_TAPS),
        ("tap_dw", C.c_int8 * TG_MAX_TAPS),
        ("partial", c_void_p), ("partial_cap", C.c_int64), ("splits", C.c_int32),
        ("blks", c_void_p), ("num_blk", C.c_int32),
        ("dtype", C.c_int32),
    ]


# name -> (restype, argtypes). Every symbol include/terragan_b200.h declares appears here; the
# CPU test-suite checks the library exports all of them.
class AdamTensor(C.Structure):
    """mirrors tg_adam_tensor"""
    _fields_ = [("param", c_void_p), ("grad", c_void_p), ("exp_avg", c_void_p), ("exp_avg_sq", c_void_p),
                ("packed_fprop", c_void_p), ("dst_fprop", c_void_p), ("packed_dgrad", c_void_p), ("dst_dgrad", c_void_p),
                ("n", C.c_int64)]


PROTOTYPES = {
    "tg_version": (c_int, []),
    "tg_last_error": (c_size_t, [C.c_char_p, c_size_t]),
    "tg_num_sms": (c_int, []),
    "tg_tmap_cache_stats": (c_int, [C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "tg_mask_window_sum": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_void_p]),
    "tg_mask_merge_up": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tg_mask_from_f32": (c_int, [c_void_p, c_long, c_void_p, c_void_p]),
    "tg_mask_to_f32": (c_int, [c_void_p, c_long, c_void_p, c_void_p]),
    "tg_conv_igemm": (c_int, [C.POINTER(ConvArgs), c_void_p]),
    "tg_conv_pool_fusable": (c_int, [C.POINTER(ConvArgs)]),
    "tg_wgrad_igemm": (c_int, [C.POINTER(WgradArgs), c_void_p]),
    "tg_wgrad_reduce": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                c_void_p]),
    "tg_wgrad_partial_floats": (C.c_int64, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "tg_bn_finalize": (c_int, [c_void_p, c_int, c_int, C.c_double, c_void_p, c_void_p, c_float, c_float,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tg_bn_eval_coeff": (c_int, [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                 c_void_p]),
    "tg_bn_apply": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_float, c_void_p,
                            c_void_p, c_void_p, c_int, c_void_p]),
    "tg_bn_bwd_reduce": (c_int, [C.POINTER(GradSrc), C.POINTER(GradSrc), c_void_p, c_int, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_int,
                                 C.POINTER(c_int), c_void_p]),
    "tg_bn_bwd_finalize": (c_int, [c_void_p, c_int, c_int, C.c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "tg_bn_bwd_apply": (c_int, [C.POINTER(GradSrc), C.POINTER(GradSrc), c_void_p, c_int, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_int, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "tg_upsample_concat": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                   c_void_p]),
    "tg_upsample_concat_bwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tg_maxpool2": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tg_maxpool2_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tg_conv_c1_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_float, c_void_p, c_int, c_void_p, c_int,
                               C.POINTER(c_int), c_void_p]),
    "tg_conv_c1_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int,
                                 c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "tg_conv_c1_wgrad_rows": (c_int, []),
    "tg_conv_to1_fwd": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, C.POINTER(c_int),
                                C.POINTER(C.c_int8), C.POINTER(C.c_int8), c_void_p, c_int, c_int, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, C.c_size_t, c_void_p]),
    "tg_conv_to1_fwd_scratch_floats": (C.c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "tg_conv_to1_fwd_kernels": (c_int, [c_int, c_int, c_int, c_int, c_int, C.POINTER(c_int), C.POINTER(C.c_int8),
                                        C.POINTER(C.c_int8), c_int, c_int]),
    "tg_adam_repack": (c_int, [C.POINTER(AdamTensor), c_int, C.c_float, C.c_float, C.c_float, C.c_float, c_int, c_void_p]),
    "tg_conv_to1_bwd_data": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, C.POINTER(C.c_int8),
                                     C.POINTER(C.c_int8), c_int, c_int, c_int, c_void_p, c_void_p]),
    "tg_conv_to1_wgrad": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                  C.POINTER(C.c_int8), C.POINTER(C.c_int8), c_void_p, c_void_p, c_int, c_void_p,
                                  c_void_p, c_int, c_void_p]),
    "tg_conv_to1_wgrad_rows": (c_int, []),
    "tg_final_bwd_pre": (c_int, [c_void_p, c_void_p, c_void_p, c_long, c_void_p, c_void_p]),
    "tg_loss_rows": (c_int, []),
    "tg_inpaint_loss_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                    c_int, c_void_p, c_void_p]),
    "tg_inpaint_loss_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p,
                                    c_void_p, c_void_p, c_void_p]),
    "tg_l1_bf16_fwd": (c_int, [c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p]),
    "tg_l1_bf16_bwd": (c_int, [c_void_p, c_void_p, c_long, c_void_p, c_int, c_void_p, c_void_p]),
    "tg_bce_logits_fwd": (c_int, [c_void_p, c_void_p, c_float, c_long, c_void_p, c_void_p]),
    "tg_bce_logits_bwd": (c_int, [c_void_p, c_void_p, c_float, c_long, c_void_p, c_void_p, c_void_p]),
    "tg_u8_prepare": (c_int, [c_void_p, c_void_p, c_long, c_void_p, c_void_p, c_void_p]),
    "tg_resize_ksize": (c_int, [c_int, c_int]),
    "tg_resize_coeffs": (c_int, [c_int, c_int, c_void_p, c_void_p, c_int]),
    "tg_resize_bilinear_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                      c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "tg_quantize_u8": (c_int, [c_void_p, c_long, c_void_p, c_void_p]),
    "tg_dsm_normalize": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "tg_morph_cross": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "tg_gauss1d_f64": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "tg_mask_shape": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, C.c_double, C.c_double, c_void_p, c_void_p]),
    "tg_quality_metrics_rows": (c_int, []),
    "tg_quality_metrics": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "tg_wgrad_partial_floats_f32": (C.c_int64, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "tg_split_tf32": (c_int, [c_void_p, c_long, c_int, c_int, c_void_p, c_void_p]),
    "tg_conv_to1_fwd_f32": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, C.POINTER(c_int),
                                    C.POINTER(C.c_int8), C.POINTER(C.c_int8), c_void_p, c_int, c_int, c_int,
                                    c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}
# fp32-storage twins with the signature of their bf16 sibling
for _bf, _f32 in (("tg_bn_apply", "tg_bn_apply_f32"), ("tg_bn_bwd_reduce", "tg_bn_bwd_reduce_f32"),
                  ("tg_bn_bwd_apply", "tg_bn_bwd_apply_f32"), ("tg_upsample_concat", "tg_upsample_concat_f32"),
                  ("tg_upsample_concat_bwd", "tg_upsample_concat_bwd_f32"), ("tg_maxpool2", "tg_maxpool2_f32"),
                  ("tg_maxpool2_bwd", "tg_maxpool2_bwd_f32"), ("tg_conv_c1_fwd", "tg_conv_c1_fwd_f32"),
                  ("tg_conv_c1_wgrad", "tg_conv_c1_wgrad_f32"), ("tg_conv_to1_bwd_data", "tg_conv_to1_bwd_data_f32"),
                  ("tg_conv_to1_wgrad", "tg_conv_to1_wgrad_f32"), ("tg_l1_bf16_fwd", "tg_l1_f32_fwd"),
                  ("tg_l1_bf16_bwd", "tg_l1_f32_bwd")):
    PROTOTYPES[_f32] = PROTOTYPES[_bf]

_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load (once) and return the kernel library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C terra-gan_b200/csrc`. There is no CPU fallback for the TERRA-GAN hot path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(l, name)  # AttributeError if the .so is stale
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().tg_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


# calls per entry point since the last reset (bench.py turns these into a kernel-launch count)
CALLS: dict = {}
# kernels launched by one successful call of each entry point
KERNELS_PER_CALL = {
    "tg_mask_window_sum": 1, "tg_mask_merge_up": 1, "tg_mask_from_f32": 1, "tg_mask_to_f32": 1,
    "tg_conv_igemm": 1, "tg_wgrad_igemm": 1, "tg_wgrad_reduce": 1,
    "tg_bn_finalize": 1, "tg_bn_eval_coeff": 1, "tg_bn_apply": 1, "tg_bn_bwd_reduce": 1, "tg_bn_bwd_finalize": 1,
    "tg_bn_bwd_apply": 1, "tg_upsample_concat": 1, "tg_upsample_concat_bwd": 1, "tg_maxpool2": 1,
    "tg_maxpool2_bwd": 1, "tg_conv_c1_fwd": 1, "tg_conv_c1_wgrad": 2, "tg_conv_to1_fwd": 1, "tg_conv_to1_fwd+tapsum": 1, "tg_conv_to1_fwd_scratch_floats": 0, "tg_conv_to1_fwd_kernels": 0,
    "tg_conv_to1_bwd_data": 1, "tg_conv_to1_wgrad": 2, "tg_final_bwd_pre": 1, "tg_inpaint_loss_fwd": 2,
    "tg_inpaint_loss_bwd": 1, "tg_l1_bf16_fwd": 2, "tg_l1_bf16_bwd": 1, "tg_bce_logits_fwd": 1, "tg_bce_logits_bwd": 1,
    "tg_quality_metrics": 2, "tg_resize_bilinear_u8": 2, "tg_dsm_normalize": 2, "tg_resize_ksize": 0, "tg_resize_coeffs": 0,
    "tg_l1_f32_fwd": 2, "tg_conv_c1_wgrad_f32": 2, "tg_conv_to1_wgrad_f32": 2, "tg_wgrad_partial_floats": 0,
    "tg_wgrad_partial_floats_f32": 0,
}


def kernel_launches() -> int:
    return sum(n * KERNELS_PER_CALL.get(k, 1) for k, n in CALLS.items())


def check(rc: int, what: str) -> None:
    CALLS[what] = CALLS.get(what, 0) + 1
    if rc != 0:
        raise RuntimeError(f"{what} failed (rc={rc}): {last_error()}")


def ptr(t) -> Optional[int]:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
