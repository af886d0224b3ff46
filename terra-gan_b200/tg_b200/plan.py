"""Host-side planning for the implicit-GEMM convolution kernels: tap tables, parity-split layout
helpers and weight packing. Pure index arithmetic (no device work) so it is unit-tested on CPU.

Layouts (see include/terragan_b200.h):
  activations   bf16 [B][P][H][W][C]; P = 1 plain NHWC, P = 4 parity-split
                (plane 2*(h&1)+(w&1) holds pixel (h>>1, w>>1))
  fprop weights bf16 [Cout][T*Cin]   column = tap*Cin + ci, tap = kh*k + kw
  dgrad weights bf16 [Cin][sum_sub T_sub*Cout]  one slab per sub-problem (input parity phase)

Reference arithmetic being re-expressed: nn.Conv2d forward / backward of
mvp_gan/src/models/pconv.py:30 and mvp_gan/src/models/discriminator.py:11.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, List, Tuple

import torch


@dataclass
class TapPlan:
    taps: List[Tuple[int, int, int]]            # (plane, dh, dw)
    subs: List[Tuple[int, int, int, int]]       # (tap_begin, tap_count, k_off [in taps], out_plane)
    kpos: List[int]                             # kernel position kh*k+kw of every tap
    in_planes: int                              # P of the tensor the taps read
    out_planes: int                             # P of the tensor written
    k: int = 0
    stride: int = 1
    pad: int = 0
    is_fprop: bool = True


def fprop_plan(k: int, stride: int, pad: int) -> TapPlan:
    """Taps of a forward conv. stride 1 reads plane 0 at (h+kh-pad, w+kw-pad); stride 2 reads the
    parity-split input: row 2*ho+kh-pad lives in plane parity (kh-pad)&1 at row ho+((kh-pad)>>1)."""
    assert stride in (1, 2)
    taps, kpos = [], []
    for kh in range(k):
        for kw in range(k):
            dh, dw = kh - pad, kw - pad
            if stride == 1:
                taps.append((0, dh, dw))
            else:
                ph, pw = dh % 2, dw % 2
                taps.append((2 * ph + pw, (dh - ph) // 2, (dw - pw) // 2))
            kpos.append(kh * k + kw)
    return TapPlan(taps, [(0, len(taps), 0, 0)], kpos, 1 if stride == 1 else 4, 1, k, stride, pad)


def dgrad_plan(k: int, stride: int, pad: int) -> TapPlan:
    """Taps of backward-data, expressed as a stride-1 gather over the output gradient g.
    stride 1: dX[h,w] = sum_{kh,kw} g[h+pad-kh, w+pad-kw] W[kh,kw].
    stride 2: input pixel (2h'+P, 2w'+Q) receives g[h'+(P+pad-kh)/2, w'+(Q+pad-kw)/2] for the kernel
    rows/cols of matching parity; one sub-problem per (P,Q) writing parity plane 2P+Q."""
    assert stride in (1, 2)
    taps, subs, kpos = [], [], []
    if stride == 1:
        for kh in range(k):
            for kw in range(k):
                taps.append((0, pad - kh, pad - kw))
                kpos.append(kh * k + kw)
        subs.append((0, len(taps), 0, 0))
        return TapPlan(taps, subs, kpos, 1, 1, k, stride, pad, False)
    for P in range(2):
        for Q in range(2):
            begin = len(taps)
            for kh in range(k):
                if (P + pad - kh) % 2:
                    continue
                for kw in range(k):
                    if (Q + pad - kw) % 2:
                        continue
                    taps.append((0, (P + pad - kh) // 2, (Q + pad - kw) // 2))
                    kpos.append(kh * k + kw)
            subs.append((begin, len(taps) - begin, begin, 2 * P + Q))
    return TapPlan(taps, subs, kpos, 1, 4, k, stride, pad, False)


def _arrange_fprop(w: torch.Tensor) -> torch.Tensor:
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)


def _arrange_dgrad(w: torch.Tensor, plan: TapPlan) -> torch.Tensor:
    co, ci, kh, kw = w.shape
    wk = w.reshape(co, ci, kh * kw)                             # [co, ci, kpos]
    idx = torch.as_tensor(plan.kpos, device=w.device, dtype=torch.long)
    sel = wk.index_select(2, idx)                               # [co, ci, T]
    return sel.permute(1, 2, 0).reshape(ci, len(plan.kpos) * co)


def pack_w_fprop(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] fp32 -> bf16 [Cout, kh*kw*Cin] (tap-major, channel-minor)."""
    return _arrange_fprop(w.detach()).to(torch.bfloat16).contiguous()


def pack_w_dgrad(w: torch.Tensor, plan: TapPlan) -> torch.Tensor:
    """[Cout, Cin, kh, kw] fp32 -> bf16 [Cin, T*Cout] in the tap order of `plan` (a dgrad plan)."""
    return _arrange_dgrad(w.detach(), plan).to(torch.bfloat16).contiguous()


def pack_w_fprop_f32(w: torch.Tensor, split=None):
    """fp32 operand matrix of the verification path. `split(w) -> (hi, lo)` (two-term TF32 split) selects the
    tf32x3 form, a (main, cross) pair: w_hi for activations x_hi, and input channels [w_hi | w_lo] for activations
    laid out [x_lo | x_hi] (the cross terms)."""
    w = w.detach().float()
    if split is None:
        return _arrange_fprop(w).contiguous()
    hi, lo = split(w)
    return _arrange_fprop(hi).contiguous(), _arrange_fprop(torch.cat([hi, lo], dim=1)).contiguous()


def pack_w_dgrad_f32(w: torch.Tensor, plan: TapPlan, split=None):
    """Same for the data-gradient operand: the contraction runs over Cout, so the split goes along dim 0."""
    w = w.detach().float()
    if split is None:
        return _arrange_dgrad(w, plan).contiguous()
    hi, lo = split(w)
    return _arrange_dgrad(hi, plan).contiguous(), _arrange_dgrad(torch.cat([hi, lo], dim=0), plan).contiguous()


def pack_scatter_index(shape, plan: Optional[TapPlan], device) -> torch.Tensor:
    """int32 [numel]: position in the packed matrix (fprop layout if plan is None, else the dgrad layout of
    `plan`) of every element of a [Cout, Cin, kh, kw] weight in its natural order. Both packings are
    permutations, so a fused optimizer step can write the bf16 copies while it updates the fp32 master."""
    n = 1
    for d in shape:
        n *= d
    src = torch.arange(n, dtype=torch.int64, device=device).reshape(shape)
    arranged = (_arrange_fprop(src) if plan is None else _arrange_dgrad(src, plan)).reshape(-1)
    dst = torch.empty(n, dtype=torch.int32, device=device)
    dst[arranged] = torch.arange(n, dtype=torch.int32, device=device)
    return dst


def to_parity_split(x: torch.Tensor) -> torch.Tensor:
    """[B, H, W, C] -> [B, 4, H/2, W/2, C] (host/test helper; kernels write this layout directly)."""
    b, h, w, c = x.shape
    return (x.reshape(b, h // 2, 2, w // 2, 2, c).permute(0, 2, 4, 1, 3, 5)
            .reshape(b, 4, h // 2, w // 2, c).contiguous())


def from_parity_split(x: torch.Tensor) -> torch.Tensor:
    """[B, 4, H/2, W/2, C] -> [B, H, W, C]."""
    b, _, h2, w2, c = x.shape
    return (x.reshape(b, 2, 2, h2, w2, c).permute(0, 3, 1, 4, 2, 5)
            .reshape(b, 2 * h2, 2 * w2, c).contiguous())


def ratio_lut(k: int) -> List[float]:
    """Mask-ratio LUT of PConv2d (pconv.py:38-40): r(s) = (k^2 / (s + 1e-8)) * [s > 0], evaluated with the
    reference's own expression on fp32 tensors (`slide_winsize / (mask_sum + 1e-8)`: torch's scalar / tensor)."""
    s = torch.arange(0, k * k + 1, dtype=torch.float32)
    r = (float(k * k) / (s + 1e-8)) * (s > 0).float()
    return [float(v) for v in r]
