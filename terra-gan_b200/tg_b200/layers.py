"""Host-side orchestration of the PConv U-Net generator, the discriminator and the VGG perceptual
branch on top of the C-ABI kernels (tg_b200.ops). Plain functions over tensors — the autograd glue
lives in tg_b200.functional, the drop-in nn.Modules in mvp_gan/src/...

Data layout inside the hot path: activations bf16 channels-last ([B,1,H,W,C] or parity-split
[B,4,H/2,W/2,C]); masks / window counts uint8; parameters fp32 masters in PyTorch layout with derived
bf16 packed copies (refreshed when the master changes: optimizer step, load_state_dict, .to()).

Reference being re-expressed: mvp_gan/src/models/{pconv,generator,discriminator}.py and the VGG
branch of mvp_gan/src/utils/losses.py — line references at each function.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch

from . import ops
from . import plan as P
from . import precision as PR
from ._lib import ACT_LEAKY, ACT_NONE, ACT_RELU

BN_EPS, BN_MOMENTUM = 1e-5, 0.1


def _tagged(tag):
    """Label every launch made inside the call with the network it belongs to (bench.py's per-network roofline)."""
    def deco(fn):
        def wrapper(*a, **k):
            prev, ops.TAG = ops.TAG, tag
            try:
                return fn(*a, **k)
            finally:
                ops.TAG = prev
        wrapper.__doc__ = fn.__doc__
        return wrapper
    return deco

class WgradLane:
    """Weight-gradient GEMMs leave the dependency chain of backward (nothing downstream reads them before the
    optimizer), so they can run on a second stream and overlap the bandwidth-bound passes (BatchNorm backward,
    up-sampling gradient) that sit between two data-gradient GEMMs. `run(fn, tensors)` orders the lane after everything
    already enqueued on the caller's stream, marks `tensors` as in use on the lane (the caching allocator must not
    recycle a gradient the lane still reads) and calls fn there; `join()` makes the caller's stream wait for the lane.
    Opt-in (TG_WGRAD_STREAM=1; otherwise fn runs inline): on the power-capped B200s of this pool the overlap lowers the
    SM clock as much as it saves (A/B at batch 64: 63.0 / 63.3 ms inline vs 63.7 / 64.3 ms with the lane, median SM clock
    1715 vs 1660 MHz) — the step is energy-bound, see DESIGN.md §6."""
    _streams: Dict[int, "torch.cuda.Stream"] = {}

    def __init__(self, device: torch.device, enabled: Optional[bool] = None):
        if enabled is None:
            enabled = os.environ.get("TG_WGRAD_STREAM", "0") == "1"
        self.stream = None
        if enabled and device.type == "cuda" and ops.PROFILE is None and not torch.cuda.is_current_stream_capturing():
            idx = device.index if device.index is not None else torch.cuda.current_device()
            if idx not in WgradLane._streams:
                WgradLane._streams[idx] = torch.cuda.Stream(device)
            self.stream = WgradLane._streams[idx]
        self.used = False

    def run(self, fn, tensors) -> None:
        if self.stream is None:
            fn()
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        for t in tensors:
            if t is not None:
                t.record_stream(self.stream)
        with torch.cuda.stream(self.stream):
            fn()
        self.used = True

    def emit(self, fn) -> None:
        """Gradient hooks (the data-parallel reducer records its 'bucket ready' event on the current stream) are called
        on the lane, after everything enqueued so far on either stream."""
        if self.stream is None or not self.used:
            fn()
            return
        self.stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self.stream):
            fn()

    def join(self) -> None:
        if self.stream is not None and self.used:
            torch.cuda.current_stream().wait_stream(self.stream)


# (name, Cin, Cout, k, stride, pad) — generator.py:13-28
ENC = [("enc1", 1, 64, 7, 2, 3), ("enc2", 64, 128, 5, 2, 2), ("enc3", 128, 256, 5, 2, 2),
       ("enc4", 256, 512, 3, 2, 1), ("enc5", 512, 512, 3, 2, 1), ("enc6", 512, 512, 3, 2, 1),
       ("enc7", 512, 512, 3, 2, 1)]
DEC = [("dec7", 1024, 512, 3, 1, 1), ("dec6", 1024, 512, 3, 1, 1), ("dec5", 1024, 512, 3, 1, 1),
       ("dec4", 768, 256, 3, 1, 1), ("dec3", 384, 128, 3, 1, 1), ("dec2", 192, 64, 3, 1, 1),
       ("dec1", 64, 64, 3, 1, 1)]


# --------------------------------------------------------------------------------------------------
# packed-weight cache
# --------------------------------------------------------------------------------------------------
class ConvPack:
    """Derived device-side state of one k x k convolution with Cin % 64 == 0: tap plans, bf16 packed
    weights for fprop / dgrad, the wgrad block table. Repacked lazily when the fp32 master changes."""

    def __init__(self, k: int, stride: int, pad: int):
        self.k, self.stride, self.pad = k, stride, pad
        self.fplan = P.fprop_plan(k, stride, pad)
        self.dplan = P.dgrad_plan(k, stride, pad)
        self.lut = P.ratio_lut(k)
        self._key = None
        self._wf = self._wd = None
        self._f32: Dict[str, tuple] = {}      # verification modes: mode -> (key, w_fprop, w_dgrad), fp32 operands
        self._blks: Dict[Tuple[int, str], torch.Tensor] = {}
        self._perm: Dict[str, torch.Tensor] = {}
        self._lut_dev: Dict[str, torch.Tensor] = {}

    def _refresh(self, w: torch.Tensor) -> None:
        key = (w.data_ptr(), w._version, str(w.device))
        if key != self._key:
            with torch.no_grad():
                self._wf = P.pack_w_fprop(w)
                self._wd = P.pack_w_dgrad(w, self.dplan)
            self._key = key

    def _refresh_f32(self, w: torch.Tensor, mode: str) -> tuple:
        key = (w.data_ptr(), w._version, str(w.device))
        ent = self._f32.get(mode)
        if ent is None or ent[0] != key:
            split = ops.split_hi_lo if mode == "tf32x3" else None
            with torch.no_grad():
                ent = (key, P.pack_w_fprop_f32(w, split), P.pack_w_dgrad_f32(w, self.dplan, split))
            self._f32[mode] = ent
        return ent

    def mark_fresh(self, w: torch.Tensor) -> None:
        """The packed copies were just written in place from `w` (tg_b200.optim.Adam): no re-pack needed."""
        self._key = (w.data_ptr(), w._version, str(w.device))

    def w_fprop(self, w: torch.Tensor, mode: str = "bf16") -> torch.Tensor:
        if mode != "bf16":
            return self._refresh_f32(w, mode)[1]
        self._refresh(w)
        return self._wf

    def w_dgrad(self, w: torch.Tensor, mode: str = "bf16") -> torch.Tensor:
        if mode != "bf16":
            return self._refresh_f32(w, mode)[2]
        self._refresh(w)
        return self._wd

    def blks(self, cin: int, device) -> torch.Tensor:
        key = (cin, str(device))
        if key not in self._blks:
            self._blks[key] = ops.wgrad_blk_table(self.fplan, cin, device)
        return self._blks[key]

    def perm(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._perm:
            self._perm[key] = torch.tensor(self.fplan.kpos, dtype=torch.int32, device=device)
        return self._perm[key]

    def lut_dev(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._lut_dev:
            self._lut_dev[key] = torch.tensor(self.lut, dtype=torch.float32, device=device)
        return self._lut_dev[key]


_MASK01 = [0.0, 1.0]
_mask01_dev: Dict[str, torch.Tensor] = {}
_index_dev: Dict[tuple, torch.Tensor] = {}


def index_dev(idx, device) -> torch.Tensor:
    """Cached device LongTensor of a constant Python index list. Indexing a CUDA tensor with a Python list uploads
    the list from pageable memory on EVERY call and synchronises the stream — two such lookups in the backward pass
    kept the host from running ahead of the GPU (a ~2 ms bubble per train step)."""
    key = (tuple(idx), str(device))
    if key not in _index_dev:
        _index_dev[key] = torch.tensor(list(idx), dtype=torch.long, device=device)
    return _index_dev[key]


def mask01_dev(device) -> torch.Tensor:
    key = str(device)
    if key not in _mask01_dev:
        _mask01_dev[key] = torch.tensor(_MASK01, dtype=torch.float32, device=device)
    return _mask01_dev[key]


# --------------------------------------------------------------------------------------------------
# BatchNorm helpers
# --------------------------------------------------------------------------------------------------
@dataclass
class BNParams:
    weight: torch.Tensor
    bias: torch.Tensor
    running_mean: torch.Tensor
    running_var: torch.Tensor
    num_batches_tracked: Optional[torch.Tensor] = None


def bn_coeffs(stats, count, bn: BNParams, training: bool):
    """(scale, shift, mean, invstd) — train: from the conv-epilogue partials (and update the running
    stats, nn.BatchNorm2d semantics momentum 0.1 / unbiased var); eval: from the running stats."""
    if training:
        out = ops.bn_finalize(stats, count, bn.weight, bn.bias, BN_EPS, BN_MOMENTUM, bn.running_mean, bn.running_var)
        if bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(1)
        return out
    scale, shift = ops.bn_eval_coeff(bn.weight, bn.bias, bn.running_mean, bn.running_var, BN_EPS)
    return scale, shift, bn.running_mean, torch.rsqrt(bn.running_var + BN_EPS)


# --------------------------------------------------------------------------------------------------
# mask pyramid of the U-Net (pconv.py:33-40 for every layer, generator.py:50-54,68-74)
# --------------------------------------------------------------------------------------------------
@dataclass
class MaskPyramid:
    m0: torch.Tensor                                  # u8 [B,H,W]
    enc_s: List[torch.Tensor] = field(default_factory=list)        # window counts of enc1..7
    enc_m: List[torch.Tensor] = field(default_factory=list)        # updated masks m1..m7
    enc_m_split: List[Optional[torch.Tensor]] = field(default_factory=list)   # m1..m6 parity-split (None for m7)
    dec_mm: List[torch.Tensor] = field(default_factory=list)       # merged input masks of dec7..dec1
    dec_s: List[torch.Tensor] = field(default_factory=list)        # window counts of dec7..dec1
    dec_m: List[torch.Tensor] = field(default_factory=list)        # updated masks of dec7..dec1


def build_mask_pyramid(m0: torch.Tensor) -> MaskPyramid:
    pyr = MaskPyramid(m0)
    m = m0
    for i, (_, _, _, k, s, p) in enumerate(ENC):
        even = (m.shape[1] // 2) % 2 == 0 and i < 6
        ssum, upd, upd_split, _ = ops.mask_window_sum(m, k, s, p, want_upd_split=even)
        pyr.enc_s.append(ssum)
        pyr.enc_m.append(upd)
        pyr.enc_m_split.append(upd_split)
        m = upd
    um = pyr.enc_m[6]
    skips = [pyr.enc_m[5], pyr.enc_m[4], pyr.enc_m[3], pyr.enc_m[2], pyr.enc_m[1], pyr.enc_m[0], m0]
    for i in range(7):
        mm = ops.mask_merge_up(um, skips[i])
        ssum, upd, _, _ = ops.mask_window_sum(mm, 3, 1, 1)
        pyr.dec_mm.append(mm)
        pyr.dec_s.append(ssum)
        pyr.dec_m.append(upd)
        um = upd
    return pyr


# --------------------------------------------------------------------------------------------------
# generator forward / backward
# --------------------------------------------------------------------------------------------------
@dataclass
class LayerSave:
    xin: Optional[torch.Tensor] = None      # conv input as consumed (masked)
    z: Optional[torch.Tensor] = None        # pre-BN conv output (bf16 [B,1,Ho,Wo,C])
    scale: Optional[torch.Tensor] = None
    shift: Optional[torch.Tensor] = None
    mean: Optional[torch.Tensor] = None
    invstd: Optional[torch.Tensor] = None


@dataclass
class GenSave:
    pyr: Optional[MaskPyramid] = None
    x: Optional[torch.Tensor] = None        # fp32 [B,H,W] input image
    layers: Dict[str, LayerSave] = field(default_factory=dict)
    y_dec1: Optional[torch.Tensor] = None
    sig: Optional[torch.Tensor] = None
    training: bool = True
    mode: str = "bf16"                      # tg_b200.precision mode of the forward pass


class GeneratorEngine:
    """PConvUNet.forward (generator.py:31-62) and its backward, hand-scheduled over the kernels."""

    def __init__(self):
        self.debug = None        # tests/tools may set a dict to capture per-layer gz tensors in backward
        self.packs = {name: ConvPack(k, s, p) for name, _, _, k, s, p in ENC + DEC}
        self.final_plan = P.fprop_plan(3, 1, 1)
        self.final_taps = [(dh, dw) for (_, dh, dw) in self.final_plan.taps]

    # ---- forward ----
    @_tagged("G")
    def forward(self, x: torch.Tensor, mask: torch.Tensor, params: Dict[str, torch.Tensor],
                bns: Dict[str, BNParams], training: bool, save: Optional[GenSave],
                trace: Optional[dict] = None) -> torch.Tensor:
        B, C1, H, W = x.shape
        if C1 != 1 or mask.shape != x.shape:
            raise RuntimeError("PConvUNet expects x and mask of shape [B,1,H,W]")
        if H % 128 or W % 128:
            raise RuntimeError(f"PConvUNet (B200 path) needs H, W divisible by 128 (7 stride-2 stages), got {H}x{W}; "
                               "the reference always feeds 512x512 tiles (train.py:68, evaluate.py:21)")
        dev = x.device
        mode = PR.get_precision()
        adt = PR.act_dtype(mode)
        x3 = x.reshape(B, H, W).contiguous().float()
        m0 = ops.mask_from_f32(mask.reshape(B, H, W).contiguous().float())
        pyr = build_mask_pyramid(m0)
        if save is not None:
            save.pyr, save.x, save.training, save.mode = pyr, x3, training, mode
        feats: List[torch.Tensor] = []      # unmasked NHWC outputs of enc1..7
        cur_split = None
        h, w = H, W
        for i, (name, cin, cout, k, s, p) in enumerate(ENC):
            pk = self.packs[name]
            ho, wo = h // 2, w // 2
            code = pyr.enc_s[i]
            if i == 0:
                w1 = params[name + ".input_conv.weight"].reshape(cout, k * k).contiguous()
                z, stats = ops.conv_c1_fwd(x3, m0, k, s, p, w1, params[name + ".input_conv.bias"], code=code,
                                           lut_dev=pk.lut_dev(dev), want_stats=training, out_dtype=adt)
                xin = None
            else:
                xin = cur_split
                z, stats = ops.conv_igemm(xin, pk.w_fprop(params[name + ".input_conv.weight"], mode), pk.fplan, (ho, wo),
                                          code=code, lut=pk.lut, bias=params[name + ".input_conv.bias"],
                                          want_stats=training)
            scale, shift, mean, invstd = bn_coeffs(stats, B * ho * wo, bns[name], training)
            want_split = i < 6
            y, ys = ops.bn_apply(z, scale, shift, ACT_RELU, 0.0, code=code, want_nhwc=True, want_split=want_split,
                                 mask_split=True)
            if save is not None:
                save.layers[name] = LayerSave(xin, z, scale, shift, mean, invstd)
            if trace is not None:
                trace[name + ".y"] = y
            feats.append(y)
            cur_split = ys
            h, w = ho, wo
        up = feats[6]
        for i, (name, cin, cout, k, s, p) in enumerate(DEC):
            pk = self.packs[name]
            skip = feats[5 - i] if i < 6 else None
            merged = ops.upsample_concat(up, skip, pyr.dec_mm[i])
            hh, ww = merged.shape[2], merged.shape[3]
            code = pyr.dec_s[i]
            if not training and save is None:
                # inference: running-statistics BatchNorm + ReLU folded into the conv epilogue (one pass less per layer)
                scale, shift, _, _ = bn_coeffs(None, B * hh * ww, bns[name], False)
                y, _ = ops.conv_igemm(merged, pk.w_fprop(params[name + ".input_conv.weight"], mode), pk.fplan, (hh, ww),
                                      code=code, lut=pk.lut, bias=params[name + ".input_conv.bias"], scale=scale,
                                      shift=shift, act=ACT_RELU)
                up = y[:, 0]
                if trace is not None:
                    trace[name + ".y"] = up
                continue
            z, stats = ops.conv_igemm(merged, pk.w_fprop(params[name + ".input_conv.weight"], mode), pk.fplan, (hh, ww),
                                      code=code, lut=pk.lut, bias=params[name + ".input_conv.bias"],
                                      want_stats=training)
            scale, shift, mean, invstd = bn_coeffs(stats, B * hh * ww, bns[name], training)
            y, _ = ops.bn_apply(z, scale, shift, ACT_RELU, 0.0, want_nhwc=True, want_split=False)
            if save is not None:
                save.layers[name] = LayerSave(merged, z, scale, shift, mean, invstd)
            if trace is not None:
                trace[name + ".y"] = y
            up = y
        wf = params["final.weight"]
        wt = wf[0].permute(1, 2, 0).reshape(9, 64).contiguous()
        out, sig = ops.conv_to1_fwd(up, False, (H, W), wt, [9], self.final_taps, params["final.bias"], (H, W), mode=1,
                                    mask=m0, xin=x3, want_sig=save is not None)
        if save is not None:
            save.y_dec1, save.sig = up, sig
        return out.reshape(B, 1, H, W)

    # ---- backward ----
    @_tagged("G")
    def backward(self, g_out: torch.Tensor, params: Dict[str, torch.Tensor], save: GenSave,
                 on_grads=None) -> Dict[str, torch.Tensor]:
        """Returns {param name: fp32 gradient}. `on_grads(names, tensors)` is called as soon as a layer's
        gradients are final (data-parallel bucketed allreduce hooks in here, tg_b200.ddp)."""
        pyr = save.pyr
        B, H, W = save.x.shape
        dev = g_out.device
        mode = save.mode
        adt, x3 = PR.act_dtype(mode), mode == "tf32x3"
        grads: Dict[str, torch.Tensor] = {}
        lane = WgradLane(dev)

        def emit(names):
            if on_grads is not None:
                lane.emit(lambda: on_grads(names, [grads[n] for n in names]))

        # final conv + sigmoid + composite — generator.py:56-62
        g_pre = ops.final_bwd_pre(g_out.reshape(B, H, W).contiguous().float(), save.sig, pyr.m0)
        wf = params["final.weight"]
        wt = wf[0].permute(1, 2, 0).reshape(9, 64).contiguous()
        grads["final.weight"] = torch.empty_like(wf)
        grads["final.bias"] = torch.empty_like(params["final.bias"])
        ops.conv_to1_wgrad(save.y_dec1, g_pre, self.final_taps, grads["final.weight"], grads["final.bias"])
        emit(["final.weight", "final.bias"])
        g_y = ops.conv_to1_bwd_data(g_pre, wt, self.final_taps, (H, W), 64, out_dtype=adt)     # grad w.r.t. dec1 output
        g_src = ops.grad_src(g_y)
        skip_grads: Dict[int, object] = {}      # encoder index -> GradSrc of the skip part
        # decoders dec1 .. dec7
        for i in range(6, -1, -1):
            name, cin, cout, k, s, p = DEC[i]
            pk = self.packs[name]
            ls = save.layers[name]
            gz, dgam, dbet, dbias = ops.bn_bwd(g_src, None, ls.z, ls.scale, ls.shift, ls.mean, ls.invstd, ACT_RELU,
                                               0.0, pyr.dec_s[i], pk.lut_dev(dev), batch_stats=save.training)
            if self.debug is not None:
                self.debug[name + ".gz"] = gz
            wkey = name + ".input_conv.weight"
            grads[wkey] = torch.empty_like(params[wkey])
            lane.run(lambda: ops.wgrad_igemm(ls.xin, gz, pk.fplan, pk.blks(cin, dev), pk.perm(dev), grads[wkey], x3=x3),
                     (ls.xin, gz, grads[wkey]))
            grads[name + ".input_conv.bias"], grads[name + ".bn.weight"], grads[name + ".bn.bias"] = dbias, dgam, dbet
            emit([wkey, name + ".input_conv.bias", name + ".bn.weight", name + ".bn.bias"])
            hh, ww = ls.xin.shape[2], ls.xin.shape[3]
            d_merged, _ = ops.conv_igemm(gz, pk.w_dgrad(params[wkey], mode), pk.dplan, (hh, ww), code=pyr.dec_mm[i],
                                         lut=_MASK01)
            cu = cin if i == 6 else cin - ENC[5 - i][2]
            if i < 6:
                skip_grads[5 - i] = ops.grad_src(d_merged, chan_off=cu)
            g_up = ops.upsample_concat_bwd(d_merged, cu)
            g_src = ops.grad_src(g_up)
        # encoders enc7 .. enc1; g_src = gradient from the decoder's up-sampling path (enc7 only)
        g_next = None    # GradSrc of dX from enc_{i+1} (parity-split)
        for i in range(6, -1, -1):
            name, cin, cout, k, s, p = ENC[i]
            pk = self.packs[name]
            ls = save.layers[name]
            if i == 6:
                a, b = g_src, None
            else:
                a, b = skip_grads[i], g_next
            gz, dgam, dbet, dbias = ops.bn_bwd(a, b, ls.z, ls.scale, ls.shift, ls.mean, ls.invstd, ACT_RELU, 0.0,
                                               pyr.enc_s[i], pk.lut_dev(dev), batch_stats=save.training)
            if self.debug is not None:
                self.debug[name + ".gz"] = gz
            wkey = name + ".input_conv.weight"
            grads[wkey] = torch.empty_like(params[wkey])
            if i == 0:
                ops.conv_c1_wgrad(save.x, pyr.m0, k, s, p, gz, False, grads[wkey], None)
            else:
                lane.run(lambda: ops.wgrad_igemm(ls.xin, gz, pk.fplan, pk.blks(cin, dev), pk.perm(dev), grads[wkey], x3=x3),
                         (ls.xin, gz, grads[wkey]))
            grads[name + ".input_conv.bias"], grads[name + ".bn.weight"], grads[name + ".bn.bias"] = dbias, dgam, dbet
            emit([wkey, name + ".input_conv.bias", name + ".bn.weight", name + ".bn.bias"])
            if i > 0:
                hi, wi = ls.xin.shape[2], ls.xin.shape[3]      # half-resolution of the layer input
                dx, _ = ops.conv_igemm(gz, pk.w_dgrad(params[wkey], mode), pk.dplan, (hi, wi),
                                       code=pyr.enc_m_split[i - 1], lut=_MASK01)
                g_next = ops.grad_src(dx, split=True)
        lane.join()
        return grads


# --------------------------------------------------------------------------------------------------
# discriminator (discriminator.py:10-26)
# --------------------------------------------------------------------------------------------------
DISC_MID = [(2, 3, 64, 128), (5, 6, 128, 256), (8, 9, 256, 512)]     # (conv idx, bn idx, Cin, Cout)


@dataclass
class DiscSave:
    img: Optional[torch.Tensor] = None
    y0: Optional[torch.Tensor] = None
    mids: List[LayerSave] = field(default_factory=list)
    y8: Optional[torch.Tensor] = None
    training: bool = True
    mode: str = "bf16"


class DiscriminatorEngine:
    def __init__(self):
        self.trace = None        # tests may set a list: every forward appends {conv idx: post-activation tensor}
        self.pack = ConvPack(4, 2, 1)
        self.p11 = P.fprop_plan(4, 1, 1)
        self.taps11 = [(dh, dw) for (_, dh, dw) in self.p11.taps]
        d0 = P.dgrad_plan(4, 2, 1)
        self.d0_plan = d0
        self.d0_taps = [(dh, dw) for (_, dh, dw) in d0.taps]
        self.d0_counts = [c for (_, c, _, _) in d0.subs]
        self._packs = {idx: ConvPack(4, 2, 1) for idx, _, _, _ in DISC_MID}

    @_tagged("D")
    def forward(self, img: torch.Tensor, params: Dict[str, torch.Tensor], bns: Dict[int, BNParams], training: bool,
                save: Optional[DiscSave]) -> torch.Tensor:
        B, C1, H, W = img.shape
        if C1 != 1:
            raise RuntimeError("Discriminator (B200 path) supports input_channels=1 (discriminator.py:7 default)")
        if H % 16 or W % 16:
            raise RuntimeError("Discriminator (B200 path) needs H, W divisible by 16")
        mode = PR.get_precision()
        x3 = img.reshape(B, H, W).contiguous().float()
        w0 = params["model.0.weight"].reshape(64, 16).contiguous()
        y0, _ = ops.conv_c1_fwd(x3, None, 4, 2, 1, w0, params["model.0.bias"], act=ACT_LEAKY, slope=0.2,
                                out_split=True, out_dtype=PR.act_dtype(mode))
        if save is not None:
            save.img, save.y0, save.training, save.mode = x3, y0, training, mode
        tr = {0: y0} if self.trace is not None else None
        cur = y0
        h, w = H // 2, W // 2
        y8 = None
        for j, (ci, bi, cin, cout) in enumerate(DISC_MID):
            pk = self._packs[ci]
            ho, wo = h // 2, w // 2
            z, stats = ops.conv_igemm(cur, pk.w_fprop(params[f"model.{ci}.weight"], mode), pk.fplan, (ho, wo),
                                      bias=params[f"model.{ci}.bias"], want_stats=training)
            scale, shift, mean, invstd = bn_coeffs(stats, B * ho * wo, bns[bi], training)
            last = j == 2
            y, ys = ops.bn_apply(z, scale, shift, ACT_LEAKY, 0.2, want_nhwc=last, want_split=not last)
            if tr is not None:
                tr[ci] = y if last else ys
            if save is not None:
                save.mids.append(LayerSave(cur, z, scale, shift, mean, invstd))
            cur = ys if not last else None
            y8 = y
            h, w = ho, wo
        w11 = params["model.11.weight"]
        wt = w11[0].permute(1, 2, 0).reshape(16, 512).contiguous()
        logits, _ = ops.conv_to1_fwd(y8, False, (h, w), wt, [16], self.taps11, params["model.11.bias"], (h - 1, w - 1))
        if save is not None:
            save.y8 = y8
        if tr is not None:
            self.trace.append(tr)
        return logits.reshape(B, 1, h - 1, w - 1)

    @_tagged("D")
    def backward(self, g_logits: torch.Tensor, params: Dict[str, torch.Tensor], save: DiscSave, need_input_grad: bool,
                 need_param_grads: bool = True, on_grads=None):
        B, H, W = save.img.shape
        dev = g_logits.device
        mode = save.mode
        adt, x3 = PR.act_dtype(mode), mode == "tf32x3"
        grads: Dict[str, torch.Tensor] = {}
        lane = WgradLane(dev)

        def emit(names):
            if on_grads is not None and need_param_grads:
                lane.emit(lambda: on_grads(names, [grads[n] for n in names]))

        g = g_logits.reshape(B, g_logits.shape[2], g_logits.shape[3]).contiguous().float()
        w11 = params["model.11.weight"]
        wt = w11[0].permute(1, 2, 0).reshape(16, 512).contiguous()
        h8, w8 = save.y8.shape[1], save.y8.shape[2]
        if need_param_grads:
            grads["model.11.weight"] = torch.empty_like(w11)
            grads["model.11.bias"] = torch.empty_like(params["model.11.bias"])
            ops.conv_to1_wgrad(save.y8, g, self.taps11, grads["model.11.weight"], grads["model.11.bias"])
            emit(["model.11.weight", "model.11.bias"])
        g_y = ops.conv_to1_bwd_data(g, wt, self.taps11, (h8, w8), 512, out_dtype=adt)
        g_src = ops.grad_src(g_y)
        for j in range(2, -1, -1):
            ci, bi, cin, cout = DISC_MID[j]
            pk = self._packs[ci]
            ls = save.mids[j]
            gz, dgam, dbet, dbias = ops.bn_bwd(g_src, None, ls.z, ls.scale, ls.shift, ls.mean, ls.invstd, ACT_LEAKY,
                                               0.2, None, None, batch_stats=save.training)
            wkey = f"model.{ci}.weight"
            if need_param_grads:
                grads[wkey] = torch.empty_like(params[wkey])
                lane.run(lambda: ops.wgrad_igemm(ls.xin, gz, pk.fplan, pk.blks(cin, dev), pk.perm(dev), grads[wkey], x3=x3),
                         (ls.xin, gz, grads[wkey]))
                grads[f"model.{ci}.bias"], grads[f"model.{bi}.weight"], grads[f"model.{bi}.bias"] = dbias, dgam, dbet
                emit([wkey, f"model.{ci}.bias", f"model.{bi}.weight", f"model.{bi}.bias"])
            hi, wi = ls.xin.shape[2], ls.xin.shape[3]
            if j > 0:
                dx, _ = ops.conv_igemm(gz, pk.w_dgrad(params[wkey], mode), pk.dplan, (hi, wi))
            else:   # into model[0]'s LeakyReLU output: fold the activation derivative (discriminator.py:14)
                dx, _ = ops.conv_igemm(gz, pk.w_dgrad(params[wkey], mode), pk.dplan, (hi, wi), gate=save.y0,
                                       gate_slope=0.2)
            g_src = ops.grad_src(dx, split=True)
            gz0 = dx
        if need_param_grads:
            grads["model.0.weight"] = torch.empty_like(params["model.0.weight"])
            grads["model.0.bias"] = torch.empty_like(params["model.0.bias"])
            ops.conv_c1_wgrad(save.img, None, 4, 2, 1, gz0, True, grads["model.0.weight"], grads["model.0.bias"])
            emit(["model.0.weight", "model.0.bias"])
        g_img = None
        if need_input_grad:
            w0 = params["model.0.weight"].reshape(64, 16)
            wt0 = w0.index_select(1, index_dev(self.d0_plan.kpos, w0.device)).t().contiguous()
            g_img, _ = ops.conv_to1_fwd(gz0, True, (H // 2, W // 2), wt0, self.d0_counts, self.d0_taps, None, (H, W))
            g_img = g_img.reshape(B, 1, H, W)
        lane.join()
        return g_img, grads


# --------------------------------------------------------------------------------------------------
# VGG16 features[:16] perceptual branch (losses.py:31-32, 79-89)
# --------------------------------------------------------------------------------------------------
VGG_CONVS = [(0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256), (14, 256, 256)]
VGG_POOL_AFTER = (2, 7)


class VggEngine:
    """Frozen VGG16[:16] on a single-channel image replicated to 3 channels (input.repeat(1,3,1,1),
    losses.py:79): conv0's three identical input channels are folded into one (weights summed)."""

    def __init__(self):
        self.trace = None        # tests may set a list: every features() call appends {conv idx: post-ReLU tensor}
        self.pack = {idx: ConvPack(3, 1, 1) for idx, _, _ in VGG_CONVS[1:]}
        d = P.dgrad_plan(3, 1, 1)
        self.d_taps = [(dh, dw) for (_, dh, dw) in d.taps]
        self.d_kpos = d.kpos
        self._w0 = None
        self._w0_key = None

    def _w0_folded(self, w0: torch.Tensor) -> torch.Tensor:
        key = (w0.data_ptr(), w0._version)
        if key != self._w0_key:
            self._w0 = w0.detach().sum(1).reshape(64, 9).contiguous().float()
            self._w0_key = key
        return self._w0

    @_tagged("VGG")
    def features(self, img: torch.Tensor, vgg: Dict[str, torch.Tensor], save: Optional[list]):
        """img fp32 [B,1,H,W] -> bf16 [B,H/4,W/4,256]; `save` collects the post-ReLU activations."""
        B, _, H, W = img.shape
        mode = PR.get_precision()
        x3 = img.reshape(B, H, W).contiguous().float()
        y, _ = ops.conv_c1_fwd(x3, None, 3, 1, 1, self._w0_folded(vgg["0.weight"]), vgg["0.bias"], act=ACT_RELU,
                               out_dtype=PR.act_dtype(mode))
        if save is not None:
            save.append(y)
        tr = {0: y} if self.trace is not None else None
        h, w = H, W
        for idx, cin, cout in VGG_CONVS[1:]:
            pk = self.pack[idx]
            if idx in VGG_POOL_AFTER:
                # conv -> ReLU -> MaxPool2d(2,2) (features[2..4], [7..9]): the pool rides in the conv epilogue; the branch
                # that needs no gradient (the target image) never writes the full-resolution activation at all
                keep = save is not None or tr is not None
                y_full, _, y = ops.conv_igemm(y, pk.w_fprop(vgg[f"{idx}.weight"], mode), pk.fplan, (h, w),
                                              bias=vgg[f"{idx}.bias"], act=ACT_RELU, pool="also" if keep else "only")
                if save is not None:
                    save.append(y_full)
                if tr is not None:
                    tr[idx] = y_full
                h, w = h // 2, w // 2
                continue
            y, _ = ops.conv_igemm(y, pk.w_fprop(vgg[f"{idx}.weight"], mode), pk.fplan, (h, w), bias=vgg[f"{idx}.bias"],
                                  act=ACT_RELU)
            if save is not None:
                save.append(y)
            if tr is not None:
                tr[idx] = y
        if tr is not None:
            self.trace.append(tr)
        return y

    @_tagged("VGG")
    def backward(self, g_feat: torch.Tensor, vgg: Dict[str, torch.Tensor], saved: list, hw,
                 mode: str = "bf16") -> torch.Tensor:
        """g_feat: gradient w.r.t. the pre-ReLU output of conv14 (bf16 [B,1,h,w,256]) -> fp32 [B,1,H,W]."""
        H, W = hw
        ys = {idx: saved[i] for i, (idx, _, _) in enumerate(VGG_CONVS)}
        gz = g_feat
        order = [14, 12, 10, 7, 5, 2]
        prev = {14: 12, 12: 10, 10: 7, 7: 5, 5: 2, 2: 0}
        for idx in order:
            pk = self.pack[idx]
            wd = pk.w_dgrad(vgg[f"{idx}.weight"], mode)
            h, w = gz.shape[2], gz.shape[3]
            pi = prev[idx]
            if pi in VGG_POOL_AFTER:     # this conv reads pool(y_prev): route through the pool, then y_prev's ReLU
                d_pool, _ = ops.conv_igemm(gz, wd, pk.dplan, (h, w))
                gz = ops.maxpool2_bwd(ys[pi][:, 0], d_pool[:, 0], relu_gate=True).unsqueeze(1)
            else:
                gz, _ = ops.conv_igemm(gz, wd, pk.dplan, (h, w), gate=ys[pi], gate_slope=0.0)
        w0 = self._w0_folded(vgg["0.weight"])
        wt0 = w0.index_select(1, index_dev(self.d_kpos, w0.device)).t().contiguous()
        g_img, _ = ops.conv_to1_fwd(gz[:, 0], False, (H, W), wt0, [9], self.d_taps, None, (H, W))
        return g_img.reshape(-1, 1, H, W)
