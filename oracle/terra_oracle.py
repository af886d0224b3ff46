"""oracle/terra_oracle.py — CPU restatement of TERRA-GAN's PConv hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this module, and only as the checker / reported baseline — never the product path (which is
hand-written sm_100a CUDA behind include/terragan_b200.h and has no CPU fallback).

What this restates (reference = FKGSOFTWARE/TERRA-GAN, paths relative to its root):
  * PConv2d.forward                      mvp_gan/src/models/pconv.py:25-50
  * PConvUNet.forward / decode_step      mvp_gan/src/models/generator.py:31-84
  * Discriminator.forward                mvp_gan/src/models/discriminator.py:10-26
  * InpaintingLoss / TV / BoundaryAware  mvp_gan/src/utils/losses.py:58-130, 386-428
  * HumanGuidedLoss                      mvp_gan/src/utils/losses.py:152-204
  * adversarial step body                mvp_gan/src/train.py:179-225
  * human-guided step body               mvp_gan/src/training/human_guided_trainer.py:101-153
The convolution / batch-norm arithmetic itself is third-party (PyTorch ATen + oneDNN, pinned by
the reference at torch==2.5.1 / torchvision==0.20.1, requirements.txt:5-6; torch 2.11.0 here), so
the restatement calls torch.nn.functional for exactly the ops the reference's nn.Modules dispatch.

It is written as pure functions over a flat `state_dict` (same keys as the reference modules), not
as nn.Modules, so it can be driven directly from checkpoints and from the CUDA modules' parameters.

Parity pinning: the reference has NO tests or golden vectors of its own. This
oracle is pinned instead against outputs of the reference's own modules run in the build container
(tests/golden/make_golden.py imports them from /root/reference by file path and writes
tests/golden/*.npz; tests/test_oracle_golden.py checks this file against those fixtures).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

# (name, Cin, Cout, k, stride, pad) — generator.py:13-28
ENC = [("enc1", 1, 64, 7, 2, 3), ("enc2", 64, 128, 5, 2, 2), ("enc3", 128, 256, 5, 2, 2),
       ("enc4", 256, 512, 3, 2, 1), ("enc5", 512, 512, 3, 2, 1), ("enc6", 512, 512, 3, 2, 1),
       ("enc7", 512, 512, 3, 2, 1)]
DEC = [("dec7", 1024, 512, 3, 1, 1), ("dec6", 1024, 512, 3, 1, 1), ("dec5", 1024, 512, 3, 1, 1),
       ("dec4", 768, 256, 3, 1, 1), ("dec3", 384, 128, 3, 1, 1), ("dec2", 192, 64, 3, 1, 1),
       ("dec1", 64, 64, 3, 1, 1)]
# discriminator.py:16-22 — (index in nn.Sequential, Cin, Cout, k, stride, pad, bn index or None)
DISC = [(0, 1, 64, 4, 2, 1, None), (2, 64, 128, 4, 2, 1, 3), (5, 128, 256, 4, 2, 1, 6),
        (8, 256, 512, 4, 2, 1, 9), (11, 512, 1, 4, 1, 1, None)]
# torchvision vgg16().features[:16]: conv indices and (Cin, Cout); max-pool after idx 3 and 8
VGG_CONVS = [(0, 3, 64), (2, 64, 64), (5, 64, 128), (7, 128, 128), (10, 128, 256), (12, 256, 256),
             (14, 256, 256)]
VGG_POOL_AFTER = (2, 7)  # a 2x2 max-pool follows the ReLU of these convs (features[4], features[9])

BN_EPS, BN_MOMENTUM = 1e-5, 0.1  # nn.BatchNorm2d defaults used by pconv.py:21 / discriminator.py:13

# --------------------------------------------------------------------------------------------------
# optional storage-rounding emulation (tests only)
# --------------------------------------------------------------------------------------------------
# The CUDA path stores activations and tensor-core weights in bf16 (fp32 accumulate). With
# `rounding(bf16_ste)` active the oracle rounds at the same points (conv weights of the tensor-core
# layers, pre-BN z, post-activation y, up-sampled features) but otherwise computes in fp32, which
# separates "the kernels compute the right thing" from "bf16 storage perturbs ill-conditioned
# quantities". With ROUND = None (default) this file is the plain fp32 restatement of the reference.
ROUND = None


class _BF16STE(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


def bf16_ste(x: Tensor) -> Tensor:
    return _BF16STE.apply(x)


class rounding:
    def __init__(self, fn):
        self.fn = fn

    def __enter__(self):
        global ROUND
        self.prev, ROUND = ROUND, self.fn
        return self

    def __exit__(self, *a):
        global ROUND
        ROUND = self.prev


def _q(x: Tensor) -> Tensor:
    return x if ROUND is None else ROUND(x)


# ---- gate tape (tests only) -----------------------------------------------------------------------
# The reference's gradients are piecewise smooth: every ReLU / LeakyReLU gate [pre > 0] and every max-pool
# routing is a discontinuity. Two correct implementations whose forward values differ by 1e-6 take different
# branches on the ~1e-6 of elements whose pre-activation lies that close to zero, and with a random upstream
# gradient that alone perturbs a weight-gradient tensor by ~sqrt(1e-6) = 1e-3 (measured: 1e-2 from the ~1e-5
# forward noise of tensor-core fp32 accumulation). To compare BACKWARD arithmetic to a sharp bound the tests
# therefore replay the branch decisions of the implementation under test into this restatement:
# `with gate_tape(gates)` makes every activation / pooling call use the recorded decision of the same call site
# (keys: "<layer>." for the PConv layers, "D<pass>.<idx>" and "vgg<pass>.<idx>", "vgg<pass>.pool<idx>"), and records how
# many decisions differed from this restatement's own and how close to the threshold those elements were.
TAPE = None


class _GatedAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pre, gate, slope):
        f = torch.where(gate, torch.ones_like(pre), torch.full_like(pre, slope))
        ctx.save_for_backward(f)
        return pre * f

    @staticmethod
    def backward(ctx, g):
        (f,) = ctx.saved_tensors
        return g * f, None, None


class gate_tape:
    def __init__(self, gates: dict):
        self.gates = gates
        self.passes: dict = {}
        self.report: dict = {}        # key -> (decisions that differ, elements, max |pre| / std(pre) among them)

    def __enter__(self):
        global TAPE
        self.prev, TAPE = TAPE, self
        return self

    def __exit__(self, *a):
        global TAPE
        TAPE = self.prev

    def next_pass(self, family: str) -> int:
        self.passes[family] = self.passes.get(family, -1) + 1
        return self.passes[family]


def _act(pre: Tensor, slope: float, key: str) -> Tensor:
    if TAPE is None or TAPE.gates.get(key) is None:
        return F.leaky_relu(pre, slope) if slope else F.relu(pre)
    gate = TAPE.gates[key]
    with torch.no_grad():
        diff = gate != (pre > 0)
        n = int(diff.sum())
        TAPE.report[key] = (n, pre.numel(), float(pre[diff].abs().max() / pre.std()) if n else 0.0)
    return _GatedAct.apply(pre, gate, slope)


class _RoutedPool(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, idx):
        ctx.save_for_backward(idx)
        ctx.shape = x.shape
        b, c, h, w = x.shape
        return x.reshape(b, c, h * w).gather(2, idx.reshape(b, c, -1)).reshape(idx.shape)

    @staticmethod
    def backward(ctx, g):
        (idx,) = ctx.saved_tensors
        b, c, h, w = ctx.shape
        gx = torch.zeros((b, c, h * w), dtype=g.dtype)
        gx.scatter_(2, idx.reshape(b, c, -1), g.reshape(b, c, -1))
        return gx.reshape(ctx.shape), None


class _SignedL1(torch.autograd.Function):
    """mean |a - b| whose backward uses a recorded sign(a - b) (the L1 kink is a branch decision too)."""
    @staticmethod
    def forward(ctx, a, b, sgn):
        ctx.save_for_backward(sgn)
        return (a - b).abs().mean()

    @staticmethod
    def backward(ctx, g):
        (sgn,) = ctx.saved_tensors
        ga = g * sgn / sgn.numel()
        return ga, -ga, None


def _l1(a: Tensor, b: Tensor, key: str) -> Tensor:
    if TAPE is None or TAPE.gates.get(key) is None:
        return F.l1_loss(a, b)
    sgn = TAPE.gates[key].to(a.dtype)
    with torch.no_grad():
        own = torch.sign(a - b)
        diff = own != sgn
        n = int(diff.sum())
        TAPE.report[key] = (n, sgn.numel(), float((a - b)[diff].abs().max() / (a - b).std()) if n else 0.0)
    return _SignedL1.apply(a, b, sgn)


class _SignedAbs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, d, sgn):
        ctx.save_for_backward(sgn)
        return d.abs()

    @staticmethod
    def backward(ctx, g):
        (sgn,) = ctx.saved_tensors
        return g * sgn, None


def _abs(d: Tensor, key: str) -> Tensor:
    """|d| of a pixel-space difference pred - target (L1, boundary and human-region terms share its sign)."""
    if TAPE is None or TAPE.gates.get(key) is None:
        return torch.abs(d)
    sgn = TAPE.gates[key].to(d.dtype)
    with torch.no_grad():
        diff = (torch.sign(d) != sgn) & (sgn != 0)
        n = int(diff.sum())
        TAPE.report[key] = (n, sgn.numel(), float(d[diff].abs().max() / d.std()) if n else 0.0)
    return _SignedAbs.apply(d, sgn * (d != 0))


def _pool(x: Tensor, key: str) -> Tensor:
    if TAPE is None or TAPE.gates.get(key) is None:
        return F.max_pool2d(x, 2, 2)
    idx = TAPE.gates[key]
    with torch.no_grad():
        _, own = F.max_pool2d(x, 2, 2, return_indices=True)
        TAPE.report[key] = (int((own != idx).sum()), idx.numel(), 0.0)
    return _RoutedPool.apply(x, idx)


# --------------------------------------------------------------------------------------------------
# deterministic parameter construction (shared by the golden generator and every parity test)
# --------------------------------------------------------------------------------------------------
def _uniform(gen: torch.Generator, shape, bound: float) -> Tensor:
    return (torch.rand(shape, generator=gen) * 2 - 1) * bound


def _conv_params(gen, sd: SD, prefix: str, cin: int, cout: int, k: int) -> None:
    bound = 1.0 / math.sqrt(cin * k * k)  # PyTorch's default conv init range
    sd[prefix + ".weight"] = _uniform(gen, (cout, cin, k, k), bound)
    sd[prefix + ".bias"] = _uniform(gen, (cout,), bound)


def _bn_params(gen, sd: SD, prefix: str, c: int) -> None:
    sd[prefix + ".weight"] = 1.0 + 0.1 * torch.randn(c, generator=gen)
    sd[prefix + ".bias"] = 0.1 * torch.randn(c, generator=gen)
    sd[prefix + ".running_mean"] = torch.zeros(c)
    sd[prefix + ".running_var"] = torch.ones(c)
    sd[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)


def make_generator_state(seed: int) -> SD:
    """state_dict with the reference PConvUNet's 114 keys/shapes (generator.py:9-29)."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for name, cin, cout, k, _, _ in ENC + DEC:
        _conv_params(gen, sd, f"{name}.input_conv", cin, cout, k)
        sd[f"{name}.mask_conv.weight"] = torch.ones(1, 1, k, k)  # pconv.py:14
        _bn_params(gen, sd, f"{name}.bn", cout)
    _conv_params(gen, sd, "final", 64, 1, 3)
    return sd


def make_discriminator_state(seed: int) -> SD:
    """state_dict with the reference Discriminator's 25 keys (discriminator.py:16-22)."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for idx, cin, cout, k, _, _, bn in DISC:
        _conv_params(gen, sd, f"model.{idx}", cin, cout, k)
        if bn is not None:
            _bn_params(gen, sd, f"model.{bn}", cout)
    return sd


def make_vgg_state(seed: int) -> SD:
    """Random-init stand-in for vgg16(IMAGENET1K_V1).features[:16] (no network in this sandbox):
    keys '<idx>.weight' / '<idx>.bias' as in torchvision's nn.Sequential."""
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for idx, cin, cout in VGG_CONVS:
        std = math.sqrt(2.0 / (cin * 9))  # He init keeps activations O(1) through the ReLUs
        sd[f"{idx}.weight"] = torch.randn((cout, cin, 3, 3), generator=gen) * std
        sd[f"{idx}.bias"] = 0.05 * torch.randn((cout,), generator=gen)
    return sd


def make_pconv_state(seed: int, cin: int, cout: int, k: int, batch_norm: bool = True) -> SD:
    gen = torch.Generator().manual_seed(seed)
    sd: SD = {}
    _conv_params(gen, sd, "input_conv", cin, cout, k)
    sd["mask_conv.weight"] = torch.ones(1, 1, k, k)
    if batch_norm:
        _bn_params(gen, sd, "bn", cout)
    return sd


# --------------------------------------------------------------------------------------------------
# synthetic inputs: DSM tiles in [0,1) and {0,1} masks, 1 = valid, 0 = hole
# --------------------------------------------------------------------------------------------------
def make_tiles(seed: int, B: int, H: int, W: Optional[int] = None) -> Tensor:
    W = H if W is None else W
    gen = torch.Generator().manual_seed(seed)
    return torch.rand((B, 1, H, W), generator=gen)


def make_mask(seed: int, B: int, H: int, kind: str = "rect", W: Optional[int] = None) -> Tensor:
    """kind: 'iid' Bernoulli(valid=.75) | 'rect' 1-4 rectangular holes | 'large' 50-80% hole |
    'ones' all valid | 'zeros' all hole."""
    W = H if W is None else W
    gen = torch.Generator().manual_seed(seed)
    if kind == "ones":
        return torch.ones(B, 1, H, W)
    if kind == "zeros":
        return torch.zeros(B, 1, H, W)
    if kind == "iid":
        return (torch.rand((B, 1, H, W), generator=gen) < 0.75).float()
    m = torch.ones(B, 1, H, W)
    for b in range(B):
        if kind == "rect":
            n = int(torch.randint(1, 5, (1,), generator=gen))
            for _ in range(n):
                hh = int(torch.randint(H // 16, H // 2 + 1, (1,), generator=gen))
                ww = int(torch.randint(W // 16, W // 2 + 1, (1,), generator=gen))
                y0 = int(torch.randint(0, H - hh + 1, (1,), generator=gen))
                x0 = int(torch.randint(0, W - ww + 1, (1,), generator=gen))
                m[b, 0, y0:y0 + hh, x0:x0 + ww] = 0
        elif kind == "large":
            hh = int(torch.randint(int(0.7 * H), int(0.9 * H) + 1, (1,), generator=gen))
            ww = int(torch.randint(int(0.7 * W), int(0.9 * W) + 1, (1,), generator=gen))
            y0 = int(torch.randint(0, H - hh + 1, (1,), generator=gen))
            x0 = int(torch.randint(0, W - ww + 1, (1,), generator=gen))
            m[b, 0, y0:y0 + hh, x0:x0 + ww] = 0
            # a few valid islands inside the hole keep the mask irregular
            for _ in range(3):
                s = max(2, H // 32)
                yy = int(torch.randint(y0, y0 + hh - s + 1, (1,), generator=gen))
                xx = int(torch.randint(x0, x0 + ww - s + 1, (1,), generator=gen))
                m[b, 0, yy:yy + s, xx:xx + s] = 1
        else:
            raise ValueError(kind)
    return m


# --------------------------------------------------------------------------------------------------
# PConv2d — pconv.py:25-50
# --------------------------------------------------------------------------------------------------
def pconv2d(x: Tensor, mask: Tensor, sd: SD, prefix: str, stride: int, pad: int, training: bool,
            batch_norm: bool = True, trace: Optional[dict] = None) -> Tuple[Tensor, Tensor]:
    """Returns (output, output_mask). BN running stats in `sd` are updated in place in training."""
    w = sd[prefix + "input_conv.weight"]
    b = sd[prefix + "input_conv.bias"]
    mw = sd[prefix + "mask_conv.weight"]
    winsize = w.shape[2] * w.shape[3]                                  # pconv.py:10
    if w.shape[1] > 1:
        w = _q(w)                                                      # (tests) bf16 tensor-core operand
    z = F.conv2d(x * mask, w, b, stride, pad)                          # :27,30  (bias inside)
    with torch.no_grad():
        msum = F.conv2d(mask, mw, None, stride, pad)                   # :34 / :38
        new_mask = (msum > 0).to(mask.dtype)                           # :35 (.float() in the reference; fp64 only in diagnostics)
        ratio = (winsize / (msum + 1e-8)) * (msum > 0).to(mask.dtype)  # :39-40 (reciprocal * k^2)
    z = _q(z * ratio)                                                  # :43
    if trace is not None:
        trace[prefix + "msum"] = msum
        trace[prefix + "z"] = z
    if batch_norm:                                                     # :46-47
        rm, rv = sd[prefix + "bn.running_mean"], sd[prefix + "bn.running_var"]
        z = F.batch_norm(z, rm, rv, sd[prefix + "bn.weight"], sd[prefix + "bn.bias"], training,
                         BN_MOMENTUM, BN_EPS)
        if training:
            sd[prefix + "bn.num_batches_tracked"] += 1
    y = _q(_act(z, 0.0, prefix))                                       # :48
    if trace is not None:
        trace[prefix + "y"] = y
        trace[prefix + "mask"] = new_mask
    return y, new_mask


# --------------------------------------------------------------------------------------------------
# PConvUNet — generator.py:31-84
# --------------------------------------------------------------------------------------------------
def _pad_to(x: Tensor, ref: Tensor) -> Tensor:                         # generator.py:78-84
    dy, dx = ref.size(2) - x.size(2), ref.size(3) - x.size(3)
    return F.pad(x, [dx // 2, dx - dx // 2, dy // 2, dy - dy // 2])


def pconv_unet(x: Tensor, mask: Tensor, sd: SD, training: bool, trace: Optional[dict] = None) -> Tensor:
    feats, masks = [], []
    h, m = x, mask
    for name, _, _, _, s, p in ENC:                                    # :33-39
        h, m = pconv2d(h, m, sd, name + ".", s, p, training, True, trace)
        feats.append(h)
        masks.append(m)
    up, um = feats[6], masks[6]
    for i, (name, _, _, _, s, p) in enumerate(DEC[:6]):                # :42-47 -> decode_step :66-76
        skip, smask = feats[5 - i], masks[5 - i]
        upf = _pad_to(_q(F.interpolate(up, scale_factor=2, mode="bilinear", align_corners=False)), skip)
        upm = _pad_to(F.interpolate(um, scale_factor=2, mode="nearest"), smask)
        merged = torch.cat([upf, skip], dim=1)                         # up-sampled channels first
        mm = torch.max(upm, smask)
        if trace is not None:
            trace[name + ".in_mask"] = mm
        up, um = pconv2d(merged, mm, sd, name + ".", s, p, training, True, trace)
    d0 = _pad_to(_q(F.interpolate(up, scale_factor=2, mode="bilinear", align_corners=False)), x)   # :50
    dm0 = _pad_to(F.interpolate(um, scale_factor=2, mode="nearest"), mask)                      # :51
    mc = torch.max(dm0, mask)                                          # :54
    if trace is not None:
        trace["dec1.in_mask"] = mc
    d0, _ = pconv2d(d0, mc, sd, "dec1.", 1, 1, training, True, trace)  # :55
    out = torch.sigmoid(F.conv2d(d0, sd["final.weight"], sd["final.bias"], 1, 1))   # :56-57
    return out * (1 - mask) + x * mask                                 # :60-62


# --------------------------------------------------------------------------------------------------
# Discriminator — discriminator.py:10-26
# --------------------------------------------------------------------------------------------------
def discriminator(img: Tensor, sd: SD, training: bool) -> Tensor:
    h = img
    pass_no = TAPE.next_pass("D") if TAPE is not None else 0
    for idx, _, _, _, s, p, bn in DISC:
        w = sd[f"model.{idx}.weight"]
        h = F.conv2d(h, _q(w) if idx in (2, 5, 8) else w, sd[f"model.{idx}.bias"], s, p)
        if idx == 11:
            break
        if bn is not None:
            h = _q(h)
        if bn is not None:
            h = F.batch_norm(h, sd[f"model.{bn}.running_mean"], sd[f"model.{bn}.running_var"],
                             sd[f"model.{bn}.weight"], sd[f"model.{bn}.bias"], training, BN_MOMENTUM, BN_EPS)
            if training:
                sd[f"model.{bn}.num_batches_tracked"] += 1
        h = _q(_act(h, 0.2, f"D{pass_no}.{idx}"))
    return h


# --------------------------------------------------------------------------------------------------
# losses — losses.py
# --------------------------------------------------------------------------------------------------
def vgg_features(x3: Tensor, vgg: SD) -> Tensor:
    """torchvision vgg16().features[:16] (conv3x3+ReLU x2, pool, x2, pool, x3), frozen, eval."""
    h = x3
    pass_no = TAPE.next_pass("vgg") if TAPE is not None else 0
    for idx, _, _ in VGG_CONVS:
        w = vgg[f"{idx}.weight"]
        h = _q(_act(F.conv2d(h, _q(w) if idx > 0 else w, vgg[f"{idx}.bias"], 1, 1), 0.0, f"vgg{pass_no}.{idx}"))
        if idx in VGG_POOL_AFTER:
            h = _pool(h, f"vgg{pass_no}.pool{idx}")
    return h


def total_variation(x: Tensor) -> Tensor:                              # losses.py:118-127
    B, _, H, W = x.shape
    count_h = x[:, :, 1:, :].numel()
    count_w = x[:, :, :, 1:].numel()
    h_tv = ((x[:, :, 1:, :] - x[:, :, :H - 1, :]) ** 2).sum()
    w_tv = ((x[:, :, :, 1:] - x[:, :, :, :W - 1]) ** 2).sum()
    return 2 * (h_tv / count_h + w_tv / count_w) / B                   # note the extra / batch


def boundary_loss(pred: Tensor, target: Tensor, mask: Tensor, eps: float = 1e-6) -> Tensor:
    """BoundaryAwareLoss.forward, losses.py:406-423."""
    dil = F.max_pool2d(mask, 3, 1, 1)
    ero = 1 - F.max_pool2d(1 - mask, 3, 1, 1)
    bd = torch.clamp(dil - ero, 0.0, 1.0)
    if bd.sum() < 1.0:                                                 # :411
        return torch.zeros((), dtype=pred.dtype)
    loss = (_abs(pred - target, "sign.pixel") * bd).sum() / (bd.sum() + eps)    # :415-416
    if torch.isnan(loss) or torch.isinf(loss):                         # :419
        return torch.zeros((), dtype=pred.dtype)
    return loss


def inpainting_loss(inp: Tensor, target: Tensor, mask: Tensor, vgg: SD, perceptual_weight: float = 0.1,
                    tv_weight: float = 0.1, boundary_weight: float = 0.5,
                    terms: Optional[dict] = None) -> Tensor:
    """InpaintingLoss.forward, losses.py:58-116."""
    l1 = _abs(inp - target, "sign.pixel").mean()                       # :73 (nn.L1Loss, mean reduction)
    total = l1
    if terms is not None:
        terms["l1"] = l1.detach()
    if perceptual_weight > 0:                                          # :77-90
        pl = _l1(vgg_features(inp.repeat(1, 3, 1, 1), vgg), vgg_features(target.repeat(1, 3, 1, 1), vgg), "l1.perceptual")
        total = total + perceptual_weight * pl
        if terms is not None:
            terms["perceptual"] = pl.detach()
    if tv_weight > 0:                                                  # :96-100
        tv = total_variation(inp * (1 - mask))
        total = total + tv_weight * tv
        if terms is not None:
            terms["tv"] = tv.detach()
    if boundary_weight > 0:                                            # :106-110
        bl = boundary_loss(inp, target, mask)
        total = total + boundary_weight * bl
        if terms is not None:
            terms["boundary"] = bl.detach()
    return total


def human_guided_loss(inp: Tensor, target: Tensor, mask: Tensor, human_mask: Optional[Tensor], vgg: SD,
                      base_w: float = 0.7, human_w: float = 0.3, boundary_weight: float = 0.5,
                      perceptual_weight: float = 0.1, tv_weight: float = 0.1,
                      terms: Optional[dict] = None) -> Tensor:
    """HumanGuidedLoss.forward, losses.py:152-204."""
    base = inpainting_loss(inp, target, mask, vgg, perceptual_weight, tv_weight, boundary_weight, terms)
    human = torch.zeros((), dtype=inp.dtype)
    if human_mask is not None:
        hm = (human_mask > 0).float()                                  # :168
        if hm.sum() > 0:                                               # :171
            human = _abs(inp * hm - target * hm, "sign.pixel").mean() if TAPE is not None else F.l1_loss(inp * hm, target * hm)   # :172-175
            if boundary_weight > 0:                                    # :178-185
                human = human + boundary_weight * boundary_loss(inp, target, hm)
    if terms is not None:
        terms["base"] = base.detach()
        terms["human"] = human.detach()
    return base_w * base + human_w * human                             # :197-200


# --------------------------------------------------------------------------------------------------
# logging-interval metrics — mvp_gan/src/utils/metrics.py:12-46, utils/experiment_tracking.py:196-231,
# mvp_gan/src/evaluation/metrics.py:79-133
# --------------------------------------------------------------------------------------------------
def quality_metrics(pred: Tensor, target: Tensor, mask: Optional[Tensor] = None) -> dict:
    mse = F.mse_loss(pred, target)                                     # metrics.py:14
    out = {"mse": mse.item(), "psnr": float("inf") if mse == 0 else (20 * torch.log10(1.0 / torch.sqrt(mse))).item()}
    C1, C2, ws = (0.01 * 1.0) ** 2, (0.03 * 1.0) ** 2, 11                # :24-39
    mu1 = F.avg_pool2d(pred, ws, stride=1, padding=ws // 2)
    mu2 = F.avg_pool2d(target, ws, stride=1, padding=ws // 2)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    s1 = F.avg_pool2d(pred * pred, ws, stride=1, padding=ws // 2) - mu1_sq
    s2 = F.avg_pool2d(target * target, ws, stride=1, padding=ws // 2) - mu2_sq
    s12 = F.avg_pool2d(pred * target, ws, stride=1, padding=ws // 2) - mu1_mu2
    out["ssim"] = (((2 * mu1_mu2 + C1) * (2 * s12 + C2)) / ((mu1_sq + mu2_sq + C1) * (s1 + s2 + C2))).mean().item()
    out["l1_distance"] = F.l1_loss(pred, target).item()                # :44-45
    out["l2_distance"] = F.mse_loss(pred, target, reduction="mean").sqrt().item()
    if mask is not None:                                               # evaluation/metrics.py:88-120
        dil = F.max_pool2d(mask, kernel_size=3, stride=1, padding=1)
        ero = 1 - F.max_pool2d(1 - mask, kernel_size=3, stride=1, padding=1)
        bd = torch.clamp(dil - ero, 0.0, 1.0)
        out["boundary_pixels"] = bd.sum().item()
        if torch.sum(bd) < 1e-6:
            out.update(boundary_mse=0.0, boundary_psnr=0.0, boundary_gradient_diff=0.0)
        else:
            bmse = torch.mean(((pred - target) * bd) ** 2)
            pd = torch.abs(pred[:, :, 1:, :] - pred[:, :, :-1, :]).mean() + torch.abs(pred[:, :, :, 1:] - pred[:, :, :, :-1]).mean()
            td = torch.abs(target[:, :, 1:, :] - target[:, :, :-1, :]).mean() + torch.abs(target[:, :, :, 1:] - target[:, :, :, :-1]).mean()
            out.update(boundary_mse=bmse.item(), boundary_psnr=(10 * torch.log10(1.0 / (bmse + 1e-6))).item(),
                       boundary_gradient_diff=torch.abs(pd - td).item())
    return out


# --------------------------------------------------------------------------------------------------
# train-step bodies
# --------------------------------------------------------------------------------------------------
def _leaf_params(sd: SD) -> List[str]:
    return [k for k, v in sd.items() if v.is_floating_point() and "running_" not in k and "mask_conv" not in k]


def _require_grad(sd: SD) -> SD:
    out = dict(sd)
    for k in _leaf_params(sd):
        out[k] = sd[k].detach().clone().requires_grad_(True)
    return out


def adam_update(p: Tensor, g: Tensor, m: Tensor, v: Tensor, step: int, lr: float, b1: float = 0.9,
                b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam defaults (train.py: Adam(lr=2e-4); human_guided_trainer.py:68 lr=1e-4), in place."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def adversarial_step(real: Tensor, masks: Tensor, g_sd: SD, d_sd: SD, vgg: SD, lr: float = 2e-4,
                     opt_state: Optional[dict] = None) -> dict:
    """One iteration of the hot loop, train.py:179-225 (boundary weight 0.5 as train() really runs).
    Mutates g_sd / d_sd (parameters after Adam, BN running stats). Returns losses and gradients."""
    G, D = _require_grad(g_sd), _require_grad(d_sd)
    masked = real * masks                                              # :181
    gen = pconv_unet(masked, masks, G, True)                           # :185
    terms: dict = {}
    g_loss = inpainting_loss(gen, real, masks, vgg, 0.1, 0.1, 0.5, terms)      # :188 (+ :110-114)
    fake_validity = discriminator(gen, D, True)                        # :202
    g_adv = F.binary_cross_entropy_with_logits(fake_validity, torch.ones_like(fake_validity))   # :203
    g_total = g_loss + g_adv                                           # :204
    g_names = _leaf_params(g_sd)
    g_grads = torch.autograd.grad(g_total, [G[k] for k in g_names])    # :206 (D grads of this pass are zeroed at :210)
    # D step — :210-219
    real_validity = discriminator(real, D, True)
    fake_validity2 = discriminator(gen.detach(), D, True)
    real_loss = F.binary_cross_entropy_with_logits(real_validity, torch.ones_like(real_validity))
    fake_loss = F.binary_cross_entropy_with_logits(fake_validity2, torch.zeros_like(fake_validity2))
    d_loss = 0.5 * (real_loss + fake_loss)
    d_names = _leaf_params(d_sd)
    d_grads = torch.autograd.grad(d_loss, [D[k] for k in d_names])
    # BN buffers were updated in place on the G / D dict copies -> carry them back
    for src, dst in ((G, g_sd), (D, d_sd)):
        for k, v in src.items():
            if "running_" in k or "num_batches" in k:
                dst[k] = v
    out = dict(gen=gen.detach(), g_loss=g_loss.detach(), g_adv=g_adv.detach(), g_total=g_total.detach(),
               d_loss=d_loss.detach(), real_loss=real_loss.detach(), fake_loss=fake_loss.detach(), terms=terms,
               g_grads=dict(zip(g_names, g_grads)), d_grads=dict(zip(d_names, d_grads)))
    if opt_state is not None:
        opt_state["step"] = opt_state.get("step", 0) + 1
        for names, grads, sd, tag in ((g_names, g_grads, g_sd, "G"), (d_names, d_grads, d_sd, "D")):
            for k, g in zip(names, grads):
                st = opt_state.setdefault(tag + k, dict(m=torch.zeros_like(g), v=torch.zeros_like(g)))
                adam_update(sd[k], g, st["m"], st["v"], opt_state["step"], lr)
    return out


def human_guided_step(images: Tensor, masks: Tensor, human_masks: Optional[Tensor], g_sd: SD, vgg: SD,
                      lr: float = 1e-4, base_w: float = 0.7, human_w: float = 0.3, boundary_weight: float = 0.5,
                      opt_state: Optional[dict] = None) -> dict:
    """human_guided_trainer.py:101-153: G -> HumanGuidedLoss -> backward -> Adam(1e-4); no discriminator."""
    G = _require_grad(g_sd)
    gen = pconv_unet(images * masks, masks, G, True)                   # :112-115
    terms: dict = {}
    loss = human_guided_loss(gen, images, masks, human_masks, vgg, base_w, human_w, boundary_weight, terms=terms)
    names = _leaf_params(g_sd)
    grads = torch.autograd.grad(loss, [G[k] for k in names])           # :151-152
    for k, v in G.items():
        if "running_" in k or "num_batches" in k:
            g_sd[k] = v
    if opt_state is not None:
        opt_state["step"] = opt_state.get("step", 0) + 1
        for k, g in zip(names, grads):
            st = opt_state.setdefault("G" + k, dict(m=torch.zeros_like(g), v=torch.zeros_like(g)))
            adam_update(g_sd[k], g, st["m"], st["v"], opt_state["step"], lr)
    return dict(gen=gen.detach(), loss=loss.detach(), terms=terms, g_grads=dict(zip(names, grads)))
