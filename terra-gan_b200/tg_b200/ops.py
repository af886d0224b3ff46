"""Thin tensor -> pointer wrappers over the C ABI (include/terragan_b200.h).

Each function validates device/dtype/contiguity, fills the ctypes argument block and launches on the
current CUDA stream. Allocation is always done here with torch (the library never allocates).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvArgs, WgradArgs, check, lib, ptr, stream_ptr
from .plan import TapPlan

BF16 = torch.bfloat16
F32 = torch.float32
_NO_POOL_FUSE = os.environ.get("TG_NO_POOL_FUSE", "") == "1"      # A/B switch: stand-alone max-pool kernels


def _act_dtype(t: torch.Tensor, name: str):
    """Activation tensors are bf16 (product path) or fp32 (verification path, tg_b200.precision)."""
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor — the TERRA-GAN hot path has no CPU fallback")
    if t.dtype not in (BF16, F32):
        raise RuntimeError(f"{name}: expected a bf16 or fp32 activation tensor, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t.dtype


def _fn(name: str, dtype):
    """Entry point of the storage type: tg_x for bf16, tg_x_f32 for fp32."""
    return getattr(lib(), name if dtype == BF16 else name + "_f32"), (name if dtype == BF16 else name + "_f32")


def split_tf32(x: torch.Tensor, layout: int) -> torch.Tensor:
    """Two-term TF32 split along the last dim (tg_split_tf32): layout 0 [hi|lo|hi], 1 [hi|hi|lo], 2 [hi|lo], 3 [hi],
    4 [lo|hi]."""
    _req(x, F32, "x")
    Cc = x.shape[-1]
    rows = x.numel() // Cc
    parts = {0: 3, 1: 3, 2: 2, 3: 1, 4: 2}[layout]
    out = torch.empty(x.shape[:-1] + (Cc * parts,), dtype=F32, device=x.device)
    check(lib().tg_split_tf32(ptr(x), rows, Cc, layout, ptr(out), stream_ptr()), "tg_split_tf32")
    return out


def split_hi_lo(w: torch.Tensor):
    """(tf32(w), tf32(w - tf32(w))) with the shape of w."""
    flat = split_tf32(w.detach().float().contiguous().reshape(1, -1), 2)
    n = w.numel()
    return flat[0, :n].reshape(w.shape), flat[0, n:].reshape(w.shape)

# bench.py sets PROFILE = [] to time every tensor-core launch with CUDA events on the launching stream:
# entries are (kind, algorithmic FLOPs, start event, end event). Events come from a pool created up front
# (profile_pool): creating a timing event costs ~100 us of host time, which made the timed region CPU-bound.
PROFILE = None
TAG = ""            # which network the engines are running ("G", "D", "VGG"): recorded with every profiled launch
_POOL: list = []


def profile_pool(n_events: int) -> None:
    """Pre-create `n_events` timing events (call before the timed region)."""
    while len(_POOL) < n_events:
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()                      # torch creates the CUDA event lazily, on the first record
        _POOL.append(ev)


def _event():
    return _POOL.pop() if _POOL else torch.cuda.Event(enable_timing=True)


def _prof_begin():
    if PROFILE is None:
        return None
    ev = _event()
    ev.record()
    return ev


def _prof_end(kind: str, flops: float, ev0, shape: str = "") -> None:
    if ev0 is None:
        return
    ev1 = _event()
    ev1.record()
    PROFILE.append((kind, flops, ev0, ev1, shape, TAG))


def _req(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor — the TERRA-GAN hot path has no CPU fallback")
    if t.device.index != torch.cuda.current_device():
        # kernels launch on the CURRENT device's stream and the library caches per-device state by cudaGetDevice()
        raise RuntimeError(f"{name}: tensor lives on cuda:{t.device.index} but the current device is "
                           f"cuda:{torch.cuda.current_device()} — call torch.cuda.set_device() (or use "
                           "`with torch.cuda.device(...)`) before running the TERRA-GAN modules on another GPU")
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
               want_stats: bool = False, out: Optional[torch.Tensor] = None,
               gate: Optional[torch.Tensor] = None, gate_slope: float = 0.0, pool: str = ""):
    """out[B][Po][Ho][Wo][N] = epilogue(sum_taps x (+) tap . w)  — see tg_conv_igemm.

    x: bf16 [B, P, H, W, C];  w_packed: bf16 [N, Ktot];  plan: fprop_plan / dgrad_plan.
    Returns (out, stats) where stats is None or fp32 [rows, 2, N] per-CTA partial sums.
    pool = "also": additionally returns the 2x2 max-pool of the output, fused into the epilogue (VGG features[4] / [9]) —
    (out, stats, pooled [B,1,Ho/2,Wo/2,N]); pool = "only": the full-resolution output is never written (out is None).
    Where the shape is not fusable (tg_conv_pool_fusable) the pool runs as its own kernel (tg_maxpool2)."""
    dt = _act_dtype(x, "x")
    addend = None
    if isinstance(w_packed, tuple):
        # tf32x3: (main, cross) operand matrices = (w_hi, [w_hi | w_lo] per tap). A first launch forms the cross terms
        # x_lo*w_hi + x_hi*w_lo on its own (small) accumulator; the main launch adds them to x_hi*w_hi before the
        # epilogue, so the tensor core's long accumulation chain carries the main term only.
        if dt != F32:
            raise RuntimeError("conv_igemm: split operands are an fp32-path feature")
        w_packed, w_cross = w_packed
        addend, _ = conv_igemm(split_tf32(x, 4), w_cross, plan, out_hw)
        x = split_tf32(x, 3)
    _req(w_packed, dt, "w_packed")
    B, P, H, W, Cc = x.shape
    N, Ktot = w_packed.shape
    Ho, Wo = out_hw
    Po = plan.out_planes
    if P != plan.in_planes:
        raise RuntimeError(f"conv_igemm: input has {P} planes, plan expects {plan.in_planes}")
    if pool not in ("", "also", "only"):
        raise ValueError("conv_igemm: pool must be '', 'also' or 'only'")
    a = ConvArgs()
    a.x, a.B, a.P, a.H, a.W, a.C = ptr(x), B, P, H, W, Cc
    a.w, a.N, a.Ktot = ptr(w_packed), N, Ktot
    a.num_sub = len(plan.subs)
    for i, (tb, tc, koff, op) in enumerate(plan.subs):
        a.sub[i].tap_begin, a.sub[i].tap_count, a.sub[i].k_off, a.sub[i].out_plane = tb, tc, koff * Cc, op
    a.num_taps = len(plan.taps)
    for i, (pl, dh, dw) in enumerate(plan.taps):
        a.tap_plane[i], a.tap_dh[i], a.tap_dw[i] = pl, dh, dw
    a.Po, a.Ho, a.Wo = Po, Ho, Wo
    a.dtype = _lib.DTYPE_BF16 if dt == BF16 else _lib.DTYPE_F32
    fuse_pool = False
    if pool:
        if gate is not None:
            raise RuntimeError("conv_igemm: pool cannot be combined with a dgrad gate")
        fuse_pool = (dt == BF16 and Po == 1 and not _NO_POOL_FUSE and bool(lib().tg_conv_pool_fusable(C.byref(a))))
    pooled = None
    if fuse_pool:
        pooled = torch.empty((B, 1, Ho // 2, Wo // 2, N), dtype=dt, device=x.device)
        a.pool_out, a.skip_out = ptr(pooled), 1 if pool == "only" else 0
    if out is None:
        if not (fuse_pool and pool == "only"):
            out = torch.empty((B, Po, Ho, Wo, N), dtype=dt, device=x.device)
    else:
        _req(out, dt, "out")
    a.out = ptr(out)
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
    a.addend = ptr(addend)
    if gate is not None:
        _req(gate, dt, "gate")
        if gate.numel() != out.numel():
            raise RuntimeError("conv_igemm: gate must have the shape of the output")
        a.gate, a.gate_slope = ptr(gate), gate_slope
    stats = None
    if want_stats:
        rows = num_sms()
        stats = torch.empty((rows, 2, N), dtype=torch.float32, device=x.device)
        a.stats, a.stats_rows_cap = ptr(stats), rows
    ev0 = _prof_begin()
    check(lib().tg_conv_igemm(C.byref(a), stream_ptr()), "tg_conv_igemm")
    _prof_end("fprop" if plan.is_fprop else "dgrad", 2.0 * B * Ho * Wo * N * len(plan.taps) * Cc, ev0,
              f"B{B} {H}x{W}x{Cc}(P{P}) -> {Ho}x{Wo}x{N}(P{Po}) taps{len(plan.taps)}")
    if stats is not None:
        stats = stats[: a.stats_rows_used]
    if pool:
        if pooled is None:                      # not fusable here: the stand-alone pooling kernel
            pooled = maxpool2(out[:, 0]).unsqueeze(1)
            if pool == "only":
                out = None
        return out, stats, pooled
    return out, stats


_BLK_DTYPE = None


def wgrad_blk_table(plan: TapPlan, Cin: int, device, gran: int = 64) -> torch.Tensor:
    """Device table of tg_wgrad_blk {int8 plane, dh, dw, pad; int32 cb; int32 row}, tap-major. `gran` = channels
    per block: 64 (bf16 storage) or 32 (fp32 storage)."""
    import numpy as np
    dt = np.dtype([("plane", "i1"), ("dh", "i1"), ("dw", "i1"), ("pad", "i1"), ("cb", "<i4"), ("row", "<i4")])
    rows = []
    for t, (pl, dh, dw) in enumerate(plan.taps):
        for cb in range(Cin // gran):
            rows.append((pl, dh, dw, 0, cb, t * Cin + cb * gran))
    arr = np.array(rows, dtype=dt)
    return torch.from_numpy(arr.view(np.uint8).reshape(-1).copy()).to(device)


_F32_BLKS: dict = {}


def wgrad_igemm(x: torch.Tensor, g: torch.Tensor, plan: TapPlan, blks: Optional[torch.Tensor],
                tap_perm: torch.Tensor, dw: torch.Tensor, accumulate: bool = False, x3: bool = False) -> None:
    """dw[N][C][kh][kw] (+)= sum_pixels x[pix (+) tap][c] * g[pix][n]   (tg_wgrad_igemm + reduce).

    x: bf16 [B, P, H, W, C] (forward input, masked); g: bf16 [B, 1, Ho, Wo, N];
    plan: the layer's *fprop* plan; dw: fp32 [N, C, k, k] contiguous.
    fp32 x / g select the verification path (kind::tf32; `blks` is built here at 32-channel granularity); with
    x3 the product is evaluated on two-term TF32 splits: [x_hi | x_lo] against [g_hi | g_lo] in one launch, of
    which the hi*hi, lo*hi and hi*lo blocks are summed."""
    dt = _act_dtype(x, "x")
    _req(g, dt, "g")
    _req(dw, torch.float32, "dw")
    if dt == F32 and x3:
        N0, C0 = g.shape[-1], x.shape[-1]
        dw2 = torch.empty((2 * N0, 2 * C0) + tuple(dw.shape[2:]), dtype=torch.float32, device=dw.device)
        wgrad_igemm(split_tf32(x, 2), split_tf32(g, 2), plan, None, tap_perm, dw2)
        tot = dw2[:N0, :C0] + dw2[:N0, C0:] + dw2[N0:, :C0]
        if accumulate:
            dw.add_(tot)
        else:
            dw.copy_(tot)
        return
    B, P, H, W, Cc = x.shape
    _, _, Ho, Wo, N = g.shape
    T = len(plan.taps)
    if dt == F32:
        key = (id(plan), Cc, str(x.device))
        if key not in _F32_BLKS:
            _F32_BLKS[key] = (plan, wgrad_blk_table(plan, Cc, x.device, 32))
        blks = _F32_BLKS[key][1]
        need = lib().tg_wgrad_partial_floats_f32(B, Ho, Wo, T, Cc, N)
    else:
        need = lib().tg_wgrad_partial_floats(B, Ho, Wo, T, Cc, N)
    partial = torch.empty((need,), dtype=torch.float32, device=x.device)
    a = WgradArgs()
    a.dtype = _lib.DTYPE_BF16 if dt == BF16 else _lib.DTYPE_F32
    a.x, a.B, a.P, a.H, a.W, a.C = ptr(x), B, P, H, W, Cc
    a.g, a.Ho, a.Wo, a.N = ptr(g), Ho, Wo, N
    a.num_taps = T
    for i, (pl, dh, dw_) in enumerate(plan.taps):
        a.tap_plane[i], a.tap_dh[i], a.tap_dw[i] = pl, dh, dw_
    a.partial, a.partial_cap = ptr(partial), need
    a.blks, a.num_blk = ptr(blks), T * (Cc // (64 if dt == BF16 else 32))
    ev0 = _prof_begin()
    check(lib().tg_wgrad_igemm(C.byref(a), stream_ptr()), "tg_wgrad_igemm")
    _prof_end("wgrad", 2.0 * B * Ho * Wo * N * T * Cc, ev0, f"B{B} x {H}x{W}x{Cc}(P{P}) g {Ho}x{Wo}x{N} taps{T} splits{a.splits}")
    check(lib().tg_wgrad_reduce(ptr(partial), a.splits, T, Cc, N, ptr(tap_perm), ptr(dw),
                                1 if accumulate else 0, stream_ptr()), "tg_wgrad_reduce")


# ------------------------------------------------------------------------------------------------
# mask pyramid
# ------------------------------------------------------------------------------------------------
def mask_from_f32(mask: torch.Tensor) -> torch.Tensor:
    _req(mask, torch.float32, "mask")
    out = torch.empty(mask.shape, dtype=torch.uint8, device=mask.device)
    check(lib().tg_mask_from_f32(ptr(mask), mask.numel(), ptr(out), stream_ptr()), "tg_mask_from_f32")
    return out


def mask_to_f32(mask: torch.Tensor) -> torch.Tensor:
    _req(mask, torch.uint8, "mask")
    out = torch.empty(mask.shape, dtype=torch.float32, device=mask.device)
    check(lib().tg_mask_to_f32(ptr(mask), mask.numel(), ptr(out), stream_ptr()), "tg_mask_to_f32")
    return out


def mask_window_sum(mask: torch.Tensor, k: int, s: int, pad: int, want_upd_split=False, want_in_split=False):
    """mask u8 [B,H,W] -> (sum u8 [B,Ho,Wo], upd u8 [B,Ho,Wo], upd_split|None, in_split|None)."""
    _req(mask, torch.uint8, "mask")
    B, H, W = mask.shape
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    dev = mask.device
    ssum = torch.empty((B, Ho, Wo), dtype=torch.uint8, device=dev)
    upd = torch.empty((B, Ho, Wo), dtype=torch.uint8, device=dev)
    upd_split = torch.empty((B, 4, Ho // 2, Wo // 2), dtype=torch.uint8, device=dev) if want_upd_split else None
    in_split = torch.empty((B, 4, H // 2, W // 2), dtype=torch.uint8, device=dev) if want_in_split else None
    check(lib().tg_mask_window_sum(ptr(mask), B, H, W, k, s, pad, ptr(ssum), ptr(upd), ptr(upd_split),
                                   ptr(in_split), stream_ptr()), "tg_mask_window_sum")
    return ssum, upd, upd_split, in_split


def mask_merge_up(up: torch.Tensor, skip: torch.Tensor) -> torch.Tensor:
    _req(up, torch.uint8, "up")
    _req(skip, torch.uint8, "skip")
    B, H, W = skip.shape
    out = torch.empty_like(skip)
    check(lib().tg_mask_merge_up(ptr(up), ptr(skip), B, H, W, ptr(out), stream_ptr()), "tg_mask_merge_up")
    return out


# ------------------------------------------------------------------------------------------------
# BatchNorm + activation
# ------------------------------------------------------------------------------------------------
def bn_finalize(partial, count, gamma, beta, eps, momentum, running_mean, running_var):
    """partial fp32 [rows,2,C] -> (scale, shift, mean, invstd); updates running stats in place."""
    rows, _, Cc = partial.shape
    dev = partial.device
    out = torch.empty((4, Cc), dtype=torch.float32, device=dev)
    check(lib().tg_bn_finalize(ptr(partial), rows, Cc, float(count), ptr(gamma), ptr(beta), eps, momentum,
                               ptr(running_mean), ptr(running_var), ptr(out[0]), ptr(out[1]), ptr(out[2]),
                               ptr(out[3]), stream_ptr()), "tg_bn_finalize")
    return out[0], out[1], out[2], out[3]


def bn_eval_coeff(gamma, beta, running_mean, running_var, eps):
    Cc = running_mean.numel()
    out = torch.empty((2, Cc), dtype=torch.float32, device=running_mean.device)
    check(lib().tg_bn_eval_coeff(Cc, ptr(gamma), ptr(beta), ptr(running_mean), ptr(running_var), eps, ptr(out[0]),
                                 ptr(out[1]), stream_ptr()), "tg_bn_eval_coeff")
    return out[0], out[1]


def bn_apply(z, scale, shift, act, slope=0.0, code=None, want_nhwc=True, want_split=False, mask_split=False):
    """z bf16 [B,1,H,W,C] or [B,H,W,C] -> (y_nhwc [B,H,W,C] | None, y_split [B,4,H/2,W/2,C] | None)."""
    dt = _act_dtype(z, "z")
    if z.dim() == 5:
        z = z[:, 0]
    B, H, W, Cc = z.shape
    y = torch.empty((B, H, W, Cc), dtype=dt, device=z.device) if want_nhwc else None
    ys = torch.empty((B, 4, H // 2, W // 2, Cc), dtype=dt, device=z.device) if want_split else None
    fn, name = _fn("tg_bn_apply", dt)
    check(fn(ptr(z), B, H, W, Cc, ptr(scale), ptr(shift), act, slope, ptr(code), ptr(y), ptr(ys),
             1 if mask_split else 0, stream_ptr()), name)
    return y, ys


def grad_src(t: Optional[torch.Tensor], chan_off: int = 0, split: bool = False) -> _lib.GradSrc:
    """Describe a gradient tensor (bf16, or fp32 on the verification path) whose last dim is the channel dim
    (pixel stride = last dim)."""
    s = _lib.GradSrc()
    if t is None:
        s.ptr, s.pix_stride, s.chan_off, s.split = None, 0, 0, 0
        s._dtype = None
    else:
        s._dtype = _act_dtype(t, "grad")
        s.ptr, s.pix_stride, s.chan_off, s.split = ptr(t), t.shape[-1], chan_off, 1 if split else 0
        s._keep = t  # the descriptor only holds a raw pointer: keep the tensor alive with it
    return s


def bn_bwd(g0: _lib.GradSrc, g1: Optional[_lib.GradSrc], z, scale, shift, mean, invstd, act, slope, code,
           lut_dev, want_dbias=True, batch_stats=True):
    """BN(+act, +mask ratio) backward. Returns (gz bf16 [B,1,H,W,C], dgamma, dbeta, dbias)."""
    dt = _act_dtype(z, "z")
    for s_ in (g0, g1):
        if s_ is not None and getattr(s_, "_dtype", dt) not in (None, dt):
            raise RuntimeError(f"bn_bwd: gradient source is {s_._dtype}, z is {dt}")
    if z.dim() == 5:
        z = z[:, 0]
    B, H, W, Cc = z.shape
    dev = z.device
    rows_cap = num_sms() * 4
    partial = torch.empty((rows_cap, 5, Cc), dtype=torch.float32, device=dev)
    used = C.c_int(0)
    g1p = C.byref(g1) if g1 is not None else None
    fn_r, name_r = _fn("tg_bn_bwd_reduce", dt)
    fn_a, name_a = _fn("tg_bn_bwd_apply", dt)
    check(fn_r(C.byref(g0), g1p, ptr(z), B, H, W, Cc, ptr(scale), ptr(shift), act, slope,
               ptr(code), ptr(lut_dev), ptr(partial), rows_cap, C.byref(used), stream_ptr()), name_r)
    outs = torch.empty((8, Cc), dtype=torch.float32, device=dev)  # coeff[5], dgamma, dbeta, dbias
    check(lib().tg_bn_bwd_finalize(ptr(partial), used.value, Cc, float(B * H * W), ptr(scale), ptr(mean),
                                   ptr(invstd), ptr(outs), ptr(outs[5]), ptr(outs[6]),
                                   ptr(outs[7]) if want_dbias else None, 0, 1 if batch_stats else 0, stream_ptr()),
          "tg_bn_bwd_finalize")
    gz = torch.empty((B, 1, H, W, Cc), dtype=dt, device=dev)
    check(fn_a(C.byref(g0), g1p, ptr(z), B, H, W, Cc, ptr(shift), ptr(outs), act, slope, ptr(code),
               ptr(lut_dev), ptr(gz), stream_ptr()), name_a)
    return gz, outs[5], outs[6], outs[7]


# ------------------------------------------------------------------------------------------------
# resampling
# ------------------------------------------------------------------------------------------------
def upsample_concat(up, skip, merged_mask):
    """up bf16 [B,h,w,Cu], skip bf16 [B,2h,2w,Cs] | None, merged_mask u8 [B,2h,2w] -> [B,1,2h,2w,Cu+Cs]."""
    dt = _act_dtype(up, "up")
    B, h, w, Cu = up.shape
    Cs = 0
    if skip is not None:
        _req(skip, dt, "skip")
        Cs = skip.shape[-1]
    out = torch.empty((B, 1, 2 * h, 2 * w, Cu + Cs), dtype=dt, device=up.device)
    fn, name = _fn("tg_upsample_concat", dt)
    check(fn(ptr(up), B, h, w, Cu, ptr(skip), Cs, ptr(merged_mask), ptr(out), stream_ptr()), name)
    return out


def upsample_concat_bwd(d_merged, Cu):
    """d_merged bf16 [B,1,2h,2w,Ctot] -> d_up bf16 [B,h,w,Cu]."""
    dt = _act_dtype(d_merged, "d_merged")
    B, _, H, W, Ct = d_merged.shape
    out = torch.empty((B, H // 2, W // 2, Cu), dtype=dt, device=d_merged.device)
    fn, name = _fn("tg_upsample_concat_bwd", dt)
    check(fn(ptr(d_merged), B, H // 2, W // 2, Cu, Ct, ptr(out), stream_ptr()), name)
    return out


def maxpool2(x):
    dt = _act_dtype(x, "x")
    B, H, W, Cc = x.shape
    y = torch.empty((B, H // 2, W // 2, Cc), dtype=dt, device=x.device)
    fn, name = _fn("tg_maxpool2", dt)
    check(fn(ptr(x), B, H, W, Cc, ptr(y), stream_ptr()), name)
    return y


def maxpool2_bwd(x, gy, relu_gate=True):
    dt = _act_dtype(x, "x")
    _req(gy, dt, "gy")
    B, H, W, Cc = x.shape
    gx = torch.empty_like(x)
    fn, name = _fn("tg_maxpool2_bwd", dt)
    check(fn(ptr(x), ptr(gy), B, H, W, Cc, 1 if relu_gate else 0, ptr(gx), stream_ptr()), name)
    return gx


# ------------------------------------------------------------------------------------------------
# bandwidth-bound convolutions
# ------------------------------------------------------------------------------------------------
def conv_c1_fwd(x, xmask, k, s, pad, wgt, bias, code=None, lut_dev=None, act=0, slope=0.0, out_split=False,
                want_stats=False, out_dtype=BF16):
    """x fp32 [B,H,W] -> bf16 (or fp32: out_dtype) [B,1,Ho,Wo,64] (or [B,4,Ho/2,Wo/2,64] if out_split), stats|None."""
    _req(x, torch.float32, "x")
    _req(wgt, torch.float32, "wgt")
    B, H, W = x.shape
    Ho, Wo = (H + 2 * pad - k) // s + 1, (W + 2 * pad - k) // s + 1
    shape = (B, 4, Ho // 2, Wo // 2, 64) if out_split else (B, 1, Ho, Wo, 64)
    out = torch.empty(shape, dtype=out_dtype, device=x.device)
    stats, used = None, C.c_int(0)
    rows = 0
    if want_stats:
        rows = num_sms() * 4
        stats = torch.empty((rows, 2, 64), dtype=torch.float32, device=x.device)
    fn, name = _fn("tg_conv_c1_fwd", out_dtype)
    ev0 = _prof_begin()
    check(fn(ptr(x), ptr(xmask), B, H, W, k, s, pad, ptr(wgt), ptr(bias), ptr(code), ptr(lut_dev),
             act, slope, ptr(out), 1 if out_split else 0, ptr(stats), rows, C.byref(used), stream_ptr()), name)
    _prof_end("thin_fprop", 2.0 * B * Ho * Wo * 64 * k * k, ev0, f"B{B} {H}x{W}x1 -> {Ho}x{Wo}x64 k{k}s{s}")
    if stats is not None:
        stats = stats[: used.value]
    return out, stats


def conv_c1_wgrad(x, xmask, k, s, pad, g, g_split, dw, db=None, accumulate=False):
    _req(x, torch.float32, "x")
    dt = _act_dtype(g, "g")
    B, H, W = x.shape
    rows = lib().tg_conv_c1_wgrad_rows()
    partial = torch.empty((rows * 64 * (k * k + 1),), dtype=torch.float32, device=x.device)
    fn, name = _fn("tg_conv_c1_wgrad", dt)
    ev0 = _prof_begin()
    check(fn(ptr(x), ptr(xmask), B, H, W, k, s, pad, ptr(g), 1 if g_split else 0, ptr(partial),
             rows, ptr(dw), ptr(db), 1 if accumulate else 0, stream_ptr()), name)
    _prof_end("thin_wgrad", 2.0 * g.numel() * k * k, ev0, f"B{B} {H}x{W}x1 k{k}s{s} g{tuple(g.shape)}")


def _tap_arrays(taps):
    n = len(taps)
    dh = (C.c_int8 * n)(*[t[0] for t in taps])
    dw = (C.c_int8 * n)(*[t[1] for t in taps])
    return dh, dw


def conv_to1_fwd(x, x_split, hw, wgt, cls_counts, taps, bias, out_hw, mode=0, mask=None, xin=None, want_sig=False):
    """x bf16 [B,H,W,C] (or parity-split of it), wgt fp32 [ntaps,C], taps [(dh,dw)] -> fp32 [B,Ho,Wo]."""
    dt = _act_dtype(x, "x")
    _req(wgt, torch.float32, "wgt")
    B = x.shape[0]
    Cc = x.shape[-1]
    H, W = hw
    Ho, Wo = out_hw
    out = torch.empty((B, Ho, Wo), dtype=torch.float32, device=x.device)
    sig = torch.empty((B, Ho, Wo), dtype=torch.float32, device=x.device) if want_sig else None
    dh, dw = _tap_arrays(taps)
    cc = (C.c_int * len(cls_counts))(*cls_counts)
    if dt == F32:
        check(lib().tg_conv_to1_fwd_f32(ptr(x), 1 if x_split else 0, B, H, W, Cc, ptr(wgt), len(cls_counts), cc, dh, dw,
                                        ptr(bias), Ho, Wo, mode, ptr(mask), ptr(xin), ptr(out), ptr(sig), stream_ptr()),
              "tg_conv_to1_fwd_f32")
        return out, sig
    one_kernel = lib().tg_conv_to1_fwd_kernels(1 if x_split else 0, H, W, Cc, len(cls_counts), cc, dh, dw, Ho, Wo) == 1
    nscr = 0 if one_kernel else lib().tg_conv_to1_fwd_scratch_floats(B, H, W, Cc, len(taps))
    scratch = torch.empty((nscr,), dtype=torch.float32, device=x.device) if nscr else None
    ev0 = _prof_begin()
    check(lib().tg_conv_to1_fwd(ptr(x), 1 if x_split else 0, B, H, W, Cc, ptr(wgt), len(cls_counts), cc, dh, dw,
                                ptr(bias), Ho, Wo, mode, ptr(mask), ptr(xin), ptr(out), ptr(sig), ptr(scratch), nscr,
                                stream_ptr()), "tg_conv_to1_fwd")
    _prof_end("thin_fprop" if mode == 1 or bias is not None else "thin_dgrad",
              2.0 * B * Ho * Wo * Cc * (len(taps) / len(cls_counts)), ev0, f"B{B} {H}x{W}x{Cc} -> {Ho}x{Wo}x1 taps{len(taps)}")
    if nscr:    # the C = 64 cases run two kernels (tap dot products, shifted sum): keep the launch count exact
        from . import _lib
        _lib.CALLS["tg_conv_to1_fwd+tapsum"] = _lib.CALLS.get("tg_conv_to1_fwd+tapsum", 0) + 1
    return out, sig


def conv_to1_bwd_data(g, wgt, taps, hw, Cc, out_dtype=BF16):
    """g fp32 [B,Ho,Wo], wgt fp32 [ntaps,C] -> dx bf16 (or fp32: out_dtype) [B,H,W,C]."""
    _req(g, torch.float32, "g")
    B, Ho, Wo = g.shape
    H, W = hw
    dx = torch.empty((B, H, W, Cc), dtype=out_dtype, device=g.device)
    dh, dw = _tap_arrays(taps)
    fn, name = _fn("tg_conv_to1_bwd_data", out_dtype)
    ev0 = _prof_begin()
    check(fn(ptr(g), B, Ho, Wo, ptr(wgt), len(taps), dh, dw, H, W, Cc, ptr(dx), stream_ptr()), name)
    _prof_end("thin_dgrad", 2.0 * B * Ho * Wo * Cc * len(taps), ev0, f"B{B} {Ho}x{Wo}x1 -> {H}x{W}x{Cc} taps{len(taps)}")
    return dx


def conv_to1_wgrad(x, g, taps, dw, db=None, accumulate=False):
    """x bf16 [B,H,W,C], g fp32 [B,Ho,Wo] -> dw fp32 [1,C,k,k] (+)=, db fp32 [1] (+)=."""
    dt = _act_dtype(x, "x")
    _req(g, torch.float32, "g")
    B, H, W, Cc = x.shape
    _, Ho, Wo = g.shape
    rows = lib().tg_conv_to1_wgrad_rows()
    partial = torch.empty((rows * len(taps) * Cc,), dtype=torch.float32, device=x.device)
    partial_b = torch.empty((rows,), dtype=torch.float32, device=x.device)
    dh, dww = _tap_arrays(taps)
    fn, name = _fn("tg_conv_to1_wgrad", dt)
    ev0 = _prof_begin()
    check(fn(ptr(x), B, H, W, Cc, ptr(g), Ho, Wo, len(taps), dh, dww, ptr(partial),
             ptr(partial_b), rows, ptr(dw), ptr(db), 1 if accumulate else 0, stream_ptr()), name)
    _prof_end("thin_wgrad", 2.0 * B * Ho * Wo * Cc * len(taps), ev0, f"B{B} x {H}x{W}x{Cc} g {Ho}x{Wo}x1 taps{len(taps)}")


def final_bwd_pre(g_out, sig, mask_u8):
    _req(g_out, torch.float32, "g_out")
    out = torch.empty_like(sig)
    check(lib().tg_final_bwd_pre(ptr(g_out), ptr(sig), ptr(mask_u8), sig.numel(), ptr(out), stream_ptr()),
          "tg_final_bwd_pre")
    return out


# ------------------------------------------------------------------------------------------------
# losses
# ------------------------------------------------------------------------------------------------
def inpaint_loss_fwd(pred, target, mask, flags=0, eps=1e-6):
    """fp32 [B,1,H,W] x3 -> terms fp32 [4] = (l1, tv, boundary, boundary_count)."""
    for n, t in (("pred", pred), ("target", target), ("mask", mask)):
        _req(t, torch.float32, n)
    if pred.dim() != 4 or pred.shape[1] != 1 or target.shape != pred.shape or mask.shape != pred.shape:
        raise RuntimeError("inpaint_loss: pred / target / mask must all be [B,1,H,W] (single-channel DSM tiles), got "
                           f"{tuple(pred.shape)}, {tuple(target.shape)}, {tuple(mask.shape)}")
    B, _, H, W = pred.shape
    rows = lib().tg_loss_rows()
    partial = torch.empty((rows * 5,), dtype=torch.float32, device=pred.device)
    terms = torch.empty((4,), dtype=torch.float32, device=pred.device)
    check(lib().tg_inpaint_loss_fwd(ptr(pred), ptr(target), ptr(mask), B, H, W, flags, eps, ptr(partial), rows,
                                    ptr(terms), stream_ptr()), "tg_inpaint_loss_fwd")
    return terms


def inpaint_loss_bwd(pred, target, mask, terms, grad_terms, flags=0, eps=1e-6):
    _req(grad_terms, torch.float32, "grad_terms")
    B, _, H, W = pred.shape
    grad = torch.empty_like(pred)
    check(lib().tg_inpaint_loss_bwd(ptr(pred), ptr(target), ptr(mask), B, H, W, flags, eps, ptr(terms),
                                    ptr(grad_terms), ptr(grad), stream_ptr()), "tg_inpaint_loss_bwd")
    return grad


def l1_bf16_fwd(a, b):
    """mean |a - b| of two activation tensors (bf16, or fp32 on the verification path)."""
    dt = _act_dtype(a, "a")
    _req(b, dt, "b")
    rows = lib().tg_loss_rows()
    partial = torch.empty((rows,), dtype=torch.float32, device=a.device)
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    fn, name = (lib().tg_l1_bf16_fwd, "tg_l1_bf16_fwd") if dt == BF16 else (lib().tg_l1_f32_fwd, "tg_l1_f32_fwd")
    check(fn(ptr(a), ptr(b), a.numel(), ptr(partial), rows, ptr(out), stream_ptr()), name)
    return out


def l1_bf16_bwd(a, b, grad_out, relu_gate=True):
    dt = _act_dtype(a, "a")
    _req(grad_out, torch.float32, "grad_out")
    ga = torch.empty_like(a)
    fn, name = (lib().tg_l1_bf16_bwd, "tg_l1_bf16_bwd") if dt == BF16 else (lib().tg_l1_f32_bwd, "tg_l1_f32_bwd")
    check(fn(ptr(a), ptr(b), a.numel(), ptr(grad_out), 1 if relu_gate else 0, ptr(ga), stream_ptr()), name)
    return ga


def bce_logits_fwd(logits, target=None, target_const=0.0):
    """mean BCE-with-logits over all elements -> fp32 [1]. target: fp32 tensor of logits' shape, or None (constant)."""
    _req(logits, torch.float32, "logits")
    if target is not None:
        _req(target, torch.float32, "target")
        if target.numel() != logits.numel():
            raise RuntimeError("bce_logits: target must have the shape of the logits")
    out = torch.empty((1,), dtype=torch.float32, device=logits.device)
    check(lib().tg_bce_logits_fwd(ptr(logits), ptr(target), float(target_const), logits.numel(), ptr(out),
                                  stream_ptr()), "tg_bce_logits_fwd")
    return out


def bce_logits_bwd(logits, grad_out, target=None, target_const=0.0):
    _req(grad_out, torch.float32, "grad_out")
    gx = torch.empty_like(logits)
    check(lib().tg_bce_logits_bwd(ptr(logits), ptr(target), float(target_const), logits.numel(), ptr(grad_out),
                                  ptr(gx), stream_ptr()), "tg_bce_logits_bwd")
    return gx


# ------------------------------------------------------------------------------------------------
# logging-interval quality metrics
# ------------------------------------------------------------------------------------------------
QUALITY_FIELDS = ("psnr", "ssim", "l1_distance", "l2_distance", "mse", "boundary_mse", "boundary_psnr",
                  "boundary_gradient_diff", "boundary_pixels")


def quality_metrics(pred, target, mask=None) -> torch.Tensor:
    """fp32 [B,1,H,W] x2 (+ mask) -> device fp32 [9] in the order of QUALITY_FIELDS (tg_quality_metrics); no host sync."""
    for n, t in (("pred", pred), ("target", target)) + ((("mask", mask),) if mask is not None else ()):
        _req(t, torch.float32, n)
    if pred.dim() != 4 or pred.shape[1] != 1 or target.shape != pred.shape or (mask is not None and mask.shape != pred.shape):
        raise RuntimeError("quality_metrics: pred / target / mask must all be [B,1,H,W] (single-channel DSM tiles)")
    B, _, H, W = pred.shape
    rows = lib().tg_quality_metrics_rows()
    partial = torch.empty((rows * 10,), dtype=torch.float64, device=pred.device)
    out = torch.empty((9,), dtype=torch.float32, device=pred.device)
    check(lib().tg_quality_metrics(ptr(pred), ptr(target), ptr(mask), B, H, W, ptr(partial), rows, ptr(out), stream_ptr()),
          "tg_quality_metrics")
    return out


# ------------------------------------------------------------------------------------------------
# batched inference I/O and DSM normalisation
# ------------------------------------------------------------------------------------------------
_RESIZE_TABLES: dict = {}


def resize_tables(in_size: int, out_size: int, device):
    """Device copies of Pillow's fixed-point bilinear coefficient tables for one axis: (bounds, kk, ksize)."""
    key = (in_size, out_size, str(device))
    if key not in _RESIZE_TABLES:
        ks = lib().tg_resize_ksize(in_size, out_size)
        if ks <= 0:
            raise RuntimeError("resize_tables: bad sizes")
        bounds = torch.empty((out_size, 2), dtype=torch.int32)
        kk = torch.empty((out_size, ks), dtype=torch.int32)
        check(lib().tg_resize_coeffs(in_size, out_size, bounds.data_ptr(), kk.data_ptr(), ks), "tg_resize_coeffs")
        _RESIZE_TABLES[key] = (bounds.to(device), kk.to(device), ks)
    return _RESIZE_TABLES[key]


def u8_prepare(img_u8: torch.Tensor, mask_u8: torch.Tensor):
    """uint8 [B,H,W] x2 -> (masked fp32 [B,1,H,W], mask fp32 [B,1,H,W]) — evaluate.py:28-33."""
    _req(img_u8, torch.uint8, "img")
    _req(mask_u8, torch.uint8, "mask")
    if img_u8.shape != mask_u8.shape or img_u8.dim() != 3:
        raise RuntimeError("u8_prepare: image and mask must both be uint8 [B,H,W]")
    B, H, W = img_u8.shape
    masked = torch.empty((B, 1, H, W), dtype=torch.float32, device=img_u8.device)
    mask = torch.empty_like(masked)
    check(lib().tg_u8_prepare(ptr(img_u8), ptr(mask_u8), img_u8.numel(), ptr(masked), ptr(mask), stream_ptr()), "tg_u8_prepare")
    return masked, mask


def resize_bilinear_u8(src: torch.Tensor, out_hw, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PIL.Image.resize(BILINEAR) of a batch of 8-bit images [B,H,W] (uint8, or fp32 quantised as (x*255).astype(uint8))."""
    if src.dtype not in (torch.uint8, torch.float32):
        raise RuntimeError("resize_bilinear_u8: src must be uint8 or fp32")
    _req(src, src.dtype, "src")
    if src.dim() == 4 and src.shape[1] == 1:
        src = src[:, 0]
    B, Hin, Win = src.shape
    Hout, Wout = out_hw
    dev = src.device
    bw, kw, ksw = resize_tables(Win, Wout, dev)
    bh, kh, ksh = resize_tables(Hin, Hout, dev)
    tmp = torch.empty((B, Hin, Wout), dtype=torch.uint8, device=dev)
    if out is None:
        out = torch.empty((B, Hout, Wout), dtype=torch.uint8, device=dev)
    check(lib().tg_resize_bilinear_u8(ptr(src), 1 if src.dtype == torch.float32 else 0, B, Hin, Win, Hout, Wout, ptr(bw), ptr(kw),
                                      ksw, ptr(bh), ptr(kh), ksh, ptr(tmp), ptr(out), stream_ptr()), "tg_resize_bilinear_u8")
    return out


def quantize_u8(x: torch.Tensor) -> torch.Tensor:
    _req(x, torch.float32, "x")
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    check(lib().tg_quantize_u8(ptr(x), x.numel(), ptr(out), stream_ptr()), "tg_quantize_u8")
    return out


def dsm_normalize(data: torch.Tensor):
    """float64 [B,H,W] (NaN = no data) -> (uint8 [B,H,W], minmax float64 [B,2]) — data_extraction.py:80-103."""
    _req(data, torch.float64, "data")
    B, H, W = data.shape
    mm = torch.empty((B, 2), dtype=torch.float64, device=data.device)
    out = torch.empty((B, H, W), dtype=torch.uint8, device=data.device)
    check(lib().tg_dsm_normalize(ptr(data), B, H, W, ptr(mm), ptr(out), stream_ptr()), "tg_dsm_normalize")
    return out, mm
