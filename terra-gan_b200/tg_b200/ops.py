"""Thin tensor -> pointer wrappers over the C ABI (include/terragan_b200.h).

Each function validates device/dtype/contiguity, fills the ctypes argument block and launches on the
current CUDA stream. Allocation is always done here with torch (the library never allocates).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvArgs, WgradArgs, check, lib, ptr, stream_ptr
from .plan import TapPlan

BF16 = torch.bfloat16


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor — the TERRA-GAN hot path has no CPU fallback")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")


def num_sms() -> int:
    n = lib().tg_num_sms()
    if n <= 0:
        raise RuntimeError("tg_num_sms failed: " + _lib.last_error())
    return n


# ------------------------------------------------------------------------------------------------
# implicit-GEMM convolution
# ------------------------------------------------------------------------------------------------
def conv_igemm(x: torch.Tensor, w_packed: torch.Tensor, plan: TapPlan, out_hw: Tuple[int, int], *,
               code: Optional[torch.Tensor] = None, lut: Optional[Sequence[float]] = None,
               bias: Optional[torch.Tensor] = None, scale: Optional[torch.Tensor] = None,
               shift: Optional[torch.Tensor] = None, act: int = 0, slope: float = 0.0,
               want_stats: bool = False, out: Optional[torch.Tensor] = None):
    """out[B][Po][Ho][Wo][N] = epilogue(sum_taps x (+) tap . w)  — see tg_conv_igemm.

    x: bf16 [B, P, H, W, C];  w_packed: bf16 [N, Ktot];  plan: fprop_plan / dgrad_plan.
    Returns (out, stats) where stats is None or fp32 [rows, 2, N] per-CTA partial sums."""
    _req(x, BF16, "x")
    _req(w_packed, BF16, "w_packed")
    B, P, H, W, Cc = x.shape
    N, Ktot = w_packed.shape
    Ho, Wo = out_hw
    Po = plan.out_planes
    if P != plan.in_planes:
        raise RuntimeError(f"conv_igemm: input has {P} planes, plan expects {plan.in_planes}")
    if out is None:
        out = torch.empty((B, Po, Ho, Wo, N), dtype=BF16, device=x.device)
    else:
        _req(out, BF16, "out")
    a = ConvArgs()
    a.x, a.B, a.P, a.H, a.W, a.C = ptr(x), B, P, H, W, Cc
    a.w, a.N, a.Ktot = ptr(w_packed), N, Ktot
    a.num_sub = len(plan.subs)
    for i, (tb, tc, koff, op) in enumerate(plan.subs):
        a.sub[i].tap_begin, a.sub[i].tap_count, a.sub[i].k_off, a.sub[i].out_plane = tb, tc, koff * Cc, op
    a.num_taps = len(plan.taps)
    for i, (pl, dh, dw) in enumerate(plan.taps):
        a.tap_plane[i], a.tap_dh[i], a.tap_dw[i] = pl, dh, dw
    a.out, a.Po, a.Ho, a.Wo = ptr(out), Po, Ho, Wo
    lut_arr = None
    if code is not None:
        _req(code, torch.uint8, "code")
        if code.numel() != B * Po * Ho * Wo:
            raise RuntimeError("conv_igemm: code must have one entry per output pixel")
        lut_arr = (C.c_float * len(lut))(*lut)
        a.code, a.lut, a.lut_len = ptr(code), lut_arr, len(lut)
    for name, t in (("bias", bias), ("scale", scale), ("shift", shift)):
        if t is not None:
            _req(t, torch.float32, name)
            if t.numel() != N:
                raise RuntimeError(f"conv_igemm: {name} must have N={N} entries")
        setattr(a, name, ptr(t))
    a.act, a.slope = act, slope
    stats = None
    if want_stats:
        rows = num_sms()
        stats = torch.empty((rows, 2, N), dtype=torch.float32, device=x.device)
        a.stats, a.stats_rows_cap = ptr(stats), rows
    check(lib().tg_conv_igemm(C.byref(a), stream_ptr()), "tg_conv_igemm")
    if stats is not None:
        stats = stats[: a.stats_rows_used]
    return out, stats


_BLK_DTYPE = None


def wgrad_blk_table(plan: TapPlan, Cin: int, device) -> torch.Tensor:
    """Device table of tg_wgrad_blk {int8 plane, dh, dw, pad; int32 cb; int32 row}, tap-major."""
    import numpy as np
    dt = np.dtype([("plane", "i1"), ("dh", "i1"), ("dw", "i1"), ("pad", "i1"), ("cb", "<i4"), ("row", "<i4")])
    rows = []
    for t, (pl, dh, dw) in enumerate(plan.taps):
        for cb in range(Cin // 64):
            rows.append((pl, dh, dw, 0, cb, t * Cin + cb * 64))
    arr = np.array(rows, dtype=dt)
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(device)


def wgrad_igemm(x: torch.Tensor, g: torch.Tensor, plan: TapPlan, blks: torch.Tensor,
                tap_perm: torch.Tensor, dw: torch.Tensor, accumulate: bool = False) -> None:
    """dw[N][C][kh][kw] (+)= sum_pixels x[pix (+) tap][c] * g[pix][n]   (tg_wgrad_igemm + reduce).

    x: bf16 [B, P, H, W, C] (forward input, masked); g: bf16 [B, 1, Ho, Wo, N];
    plan: the layer's *fprop* plan; dw: fp32 [N, C, k, k] contiguous."""
    _req(x, BF16, "x")
    _req(g, BF16, "g")
    _req(dw, torch.float32, "dw")
    B, P, H, W, Cc = x.shape
    _, _, Ho, Wo, N = g.shape
    T = len(plan.taps)
    need = lib().tg_wgrad_partial_floats(B, Ho, Wo, T, Cc, N)
    partial = torch.empty((need,), dtype=torch.float32, device=x.device)
    a = WgradArgs()
    a.x, a.B, a.P, a.H, a.W, a.C = ptr(x), B, P, H, W, Cc
    a.g, a.Ho, a.Wo, a.N = ptr(g), Ho, Wo, N
    a.num_taps = T
    for i, (pl, dh, dw_) in enumerate(plan.taps):
        a.tap_plane[i], a.tap_dh[i], a.tap_dw[i] = pl, dh, dw_
    a.partial, a.partial_cap = ptr(partial), need
    a.blks, a.num_blk = ptr(blks), T * (Cc // 64)
    check(lib().tg_wgrad_igemm(C.byref(a), stream_ptr()), "tg_wgrad_igemm")
    check(lib().tg_wgrad_reduce(ptr(partial), a.splits, T, Cc, N, ptr(tap_perm), ptr(dw),
                                1 if accumulate else 0, stream_ptr()), "tg_wgrad_reduce")
