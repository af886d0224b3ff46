"""torch.autograd.Function glue between the drop-in nn.Modules (mvp_gan/src/...) and the engines in
tg_b200.layers. Each Function takes/returns ordinary fp32 NCHW tensors (the reference's module API)
and keeps the bf16 channels-last hot-path tensors private to the engine.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch

from . import ops
from . import plan as P
from . import precision as PR
from ._lib import ACT_NONE, ACT_RELU
from .layers import (BN_EPS, BNParams, ConvPack, DiscSave, DiscriminatorEngine, GenSave, GeneratorEngine,
                     VggEngine, bn_coeffs)

# A data-parallel wrapper installs a callback here: on_grads(owner_module, names, tensors) is invoked
# from inside backward as soon as a layer's gradients are final (tg_b200.ddp).
_grad_hooks: Dict[int, object] = {}


def set_grad_hook(module, fn) -> None:
    if fn is None:
        _grad_hooks.pop(id(module), None)
    else:
        _grad_hooks[id(module)] = fn


def _hook_for(module):
    fn = _grad_hooks.get(id(module))
    if fn is None:
        return None
    return lambda names, tensors: fn(module, names, tensors)


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: input is on {t.device}; the B200 TERRA-GAN path runs hand-written sm_100a CUDA "
                           "kernels only and has no CPU fallback (move the module and its inputs to a CUDA device)")


def _check_versions(ctx, what: str) -> None:
    """The weights are read again in backward (packed dgrad operands): autograd's own saved-tensor version check does
    not see them, so an optimizer step between forward and backward is caught here instead of silently using the
    new weights."""
    for name, p, v in zip(ctx.names, ctx.params.values(), ctx.versions):
        if p._version != v:
            raise RuntimeError(f"{what}: parameter {name} was modified in place between forward and backward "
                               f"(version {v} -> {p._version}); one of the variables needed for gradient computation "
                               "has been modified by an inplace operation")


class GeneratorFn(torch.autograd.Function):
    """PConvUNet.forward — generator.py:31-62."""

    @staticmethod
    def forward(ctx, x, mask, module, names, *plist):
        _require_cuda(x, "PConvUNet")
        params = dict(zip(names, plist))
        need = any(ctx.needs_input_grad)      # grad mode is already off inside Function.forward
        save = GenSave() if need else None
        with torch.no_grad():
            out = module._engine.forward(x, mask, params, module._bn_params(), module.training, save,
                                         getattr(module, "_trace", None))
        ctx.module, ctx.names, ctx.save = module, names, save
        ctx.params = params
        ctx.versions = [p._version for p in plist]
        return out

    @staticmethod
    def backward(ctx, g_out):
        if ctx.save is None:
            raise RuntimeError("PConvUNet: backward called but no state was saved")
        _check_versions(ctx, "PConvUNet")
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("PConvUNet (B200 path): gradient w.r.t. the input image is not implemented "
                                      "(no caller on the reference path needs it: train.py:181-185)")
        with torch.no_grad():
            grads = ctx.module._engine.backward(g_out.contiguous(), ctx.params, ctx.save, _hook_for(ctx.module))
        ctx.save = None
        return (None, None, None, None) + tuple(grads.get(n) for n in ctx.names)


class DiscriminatorFn(torch.autograd.Function):
    """Discriminator.forward — discriminator.py:25-26."""

    @staticmethod
    def forward(ctx, img, module, names, *plist):
        _require_cuda(img, "Discriminator")
        params = dict(zip(names, plist))
        need = any(ctx.needs_input_grad)
        save = DiscSave() if need else None
        with torch.no_grad():
            out = module._engine.forward(img, params, module._bn_params(), module.training, save)
        ctx.module, ctx.names, ctx.save, ctx.params = module, names, save, params
        ctx.versions = [p._version for p in plist]
        return out

    @staticmethod
    def backward(ctx, g):
        _check_versions(ctx, "Discriminator")
        need_p = any(ctx.needs_input_grad[3:])
        with torch.no_grad():
            g_img, grads = ctx.module._engine.backward(g.contiguous(), ctx.params, ctx.save, ctx.needs_input_grad[0],
                                                       need_p, _hook_for(ctx.module))
        ctx.save = None
        return (g_img, None, None) + tuple(grads.get(n) if ctx.needs_input_grad[3 + i] else None
                                           for i, n in enumerate(ctx.names))


class PerceptualFn(torch.autograd.Function):
    """L1(vgg(input.repeat(1,3,1,1)), vgg(target.repeat(1,3,1,1))) — losses.py:79-89. VGG is frozen."""

    @staticmethod
    def forward(ctx, inp, target, engine: VggEngine, vgg: Dict[str, torch.Tensor]):
        _require_cuda(inp, "InpaintingLoss")
        need = ctx.needs_input_grad[0]
        saved: Optional[list] = [] if need else None
        with torch.no_grad():
            fi = engine.features(inp, vgg, saved)
            ft = engine.features(target, vgg, None)
            loss = ops.l1_bf16_fwd(fi, ft)
        ctx.engine, ctx.vgg, ctx.saved, ctx.fi, ctx.ft = engine, vgg, saved, fi, ft
        ctx.hw = (inp.shape[2], inp.shape[3])
        ctx.mode = PR.get_precision()
        return loss.reshape(())

    @staticmethod
    def backward(ctx, go):
        with torch.no_grad():
            gfeat = ops.l1_bf16_bwd(ctx.fi, ctx.ft, go.reshape(1).float().contiguous(), relu_gate=True)
            g_img = ctx.engine.backward(gfeat, ctx.vgg, ctx.saved, ctx.hw, ctx.mode)
        ctx.saved = ctx.fi = ctx.ft = None
        return g_img, None, None, None


class InpaintTermsFn(torch.autograd.Function):
    """(L1, TV(pred*(1-mask)), boundary loss) in one fused pass — losses.py:73, 96-100/118-127, 406-423."""

    @staticmethod
    def forward(ctx, pred, target, mask, flags: int, eps: float):
        _require_cuda(pred, "InpaintingLoss")
        p = pred.detach().float().contiguous()
        t = target.detach().float().contiguous()
        m = mask.detach().float().contiguous()
        terms = ops.inpaint_loss_fwd(p, t, m, flags, eps)
        ctx.saved = (p, t, m, terms)
        ctx.flags, ctx.eps = flags, eps
        return terms[:3].clone()

    @staticmethod
    def backward(ctx, g_terms):
        p, t, m, terms = ctx.saved
        grad = ops.inpaint_loss_bwd(p, t, m, terms, g_terms.float().contiguous(), ctx.flags, ctx.eps)
        ctx.saved = None
        return grad, None, None, None, None


class BCEWithLogitsFn(torch.autograd.Function):
    """nn.BCEWithLogitsLoss()(logits, target) with mean reduction — train.py:115,203,215-216. `target` is a tensor or a
    Python float (the loops pass ones_like / zeros_like: a constant needs no target tensor at all)."""

    @staticmethod
    def forward(ctx, logits, target):
        _require_cuda(logits, "BCEWithLogits")
        x = logits.detach().float().contiguous()
        if isinstance(target, torch.Tensor):
            t, tc = target.detach().float().contiguous(), 0.0
        else:
            t, tc = None, float(target)
        out = ops.bce_logits_fwd(x, t, tc)
        ctx.saved = (x, t, tc, logits.shape)
        return out.reshape(())

    @staticmethod
    def backward(ctx, go):
        x, t, tc, shape = ctx.saved
        gx = ops.bce_logits_bwd(x, go.reshape(1).float().contiguous(), t, tc)
        ctx.saved = None
        return gx.reshape(shape), None


class BCEWithLogitsLoss(torch.nn.Module):
    """Drop-in for the `adversarial_loss = nn.BCEWithLogitsLoss()` object of train.py:115 on the B200 path: same call
    signature `(input, target) -> scalar`; targets that are constant tensors may also be passed as a Python float."""

    def forward(self, input, target):
        return BCEWithLogitsFn.apply(input, target)


class PConv2dFn(torch.autograd.Function):
    """Stand-alone PConv2d.forward(input, mask) -> (output, output_mask) — pconv.py:25-50 — for shapes
    with Cin == 1 (Cout == 64; 7x7/s2, 4x4/s2, 3x3/s1) or Cin % 64 == 0 (stride 1 or 2)."""

    @staticmethod
    def forward(ctx, x, mask, module, weight, bias, gamma, beta):
        _require_cuda(x, "PConv2d")
        k, s, p = module._k, module._stride, module._pad
        B, cin, H, W = x.shape
        cout = weight.shape[0]
        pk: ConvPack = module._pack
        dev = x.device
        mode = PR.get_precision()
        adt = PR.act_dtype(mode)
        with torch.no_grad():
            m8 = ops.mask_from_f32(mask.reshape(B, H, W).contiguous().float())
            ssum, upd, _, m_split = ops.mask_window_sum(m8, k, s, p, want_in_split=(s == 2 and cin > 1))
            ho, wo = ssum.shape[1], ssum.shape[2]
            training = module.training
            want_stats = module.batch_norm and training
            epi_act = ACT_NONE if module.batch_norm else ACT_RELU
            if cin == 1:
                if cout != 64 or (k, s) not in ((7, 2), (4, 2), (3, 1)):
                    raise NotImplementedError(f"PConv2d (B200 path): Cin=1 supports Cout=64 with k/s in 7/2, 4/2, 3/1; "
                                              f"got Cout={cout}, k={k}, s={s}")
                x3 = x.reshape(B, H, W).contiguous().float()
                z, stats = ops.conv_c1_fwd(x3, m8, k, s, p, weight.reshape(cout, k * k).contiguous(), bias, code=ssum,
                                           lut_dev=pk.lut_dev(dev), act=epi_act, want_stats=want_stats, out_dtype=adt)
                xin = x3
            else:
                if cin % 64 or cout % 64 or s not in (1, 2) or (s == 2 and (H % 2 or W % 2)):
                    raise NotImplementedError("PConv2d (B200 path): channels must be multiples of 64 and stride 1 or 2 "
                                              f"(even H, W); got Cin={cin}, Cout={cout}, stride={s}, {H}x{W}")
                xm = (x * mask).permute(0, 2, 3, 1).to(adt).contiguous()
                xin = P.to_parity_split(xm) if s == 2 else xm.unsqueeze(1)
                z, stats = ops.conv_igemm(xin, pk.w_fprop(weight, mode), pk.fplan, (ho, wo), code=ssum, lut=pk.lut, bias=bias,
                                          act=epi_act, want_stats=want_stats)
            if module.batch_norm:
                bn = module.bn
                scale, shift, mean, invstd = bn_coeffs(stats, B * ho * wo, BNParams(gamma, beta, bn.running_mean,
                                                                                   bn.running_var,
                                                                                   bn.num_batches_tracked), training)
                y, _ = ops.bn_apply(z, scale, shift, ACT_RELU)
            else:
                scale = torch.ones(cout, device=dev)
                shift = torch.zeros(cout, device=dev)
                mean, invstd = shift, scale
                y = z[:, 0]
            out = y.permute(0, 3, 1, 2).float()
            out_mask = ops.mask_to_f32(upd).reshape(B, 1, ho, wo)
        ctx.state = (module, xin, z, scale, shift, mean, invstd, ssum, m8, m_split, (B, cin, H, W), weight)
        ctx.mode = mode
        ctx.mark_non_differentiable(out_mask)
        return out, out_mask

    @staticmethod
    def backward(ctx, g_out, _g_mask):
        module, xin, z, scale, shift, mean, invstd, ssum, m8, m_split, shape, weight = ctx.state
        B, cin, H, W = shape
        k, s, p = module._k, module._stride, module._pad
        pk: ConvPack = module._pack
        dev = g_out.device
        cout = weight.shape[0]
        mode = ctx.mode
        with torch.no_grad():
            g = g_out.permute(0, 2, 3, 1).to(PR.act_dtype(mode)).contiguous()
            gz, dgam, dbet, dbias = ops.bn_bwd(ops.grad_src(g), None, z, scale, shift, mean, invstd, ACT_RELU, 0.0, ssum,
                                               pk.lut_dev(dev), batch_stats=module.batch_norm and module.training)
            dw = torch.empty_like(weight)
            gx = None
            if cin == 1:
                ops.conv_c1_wgrad(xin, m8, k, s, p, gz, False, dw, None)
                if ctx.needs_input_grad[0]:
                    dpl = pk.dplan
                    from .layers import index_dev
                    wt = weight.reshape(cout, k * k).index_select(1, index_dev(dpl.kpos, weight.device)).t().contiguous()
                    taps = [(dh, dw_) for (_, dh, dw_) in dpl.taps]
                    counts = [c for (_, c, _, _) in dpl.subs]
                    ho, wo = ssum.shape[1], ssum.shape[2]
                    gx, _ = ops.conv_to1_fwd(gz[:, 0], False, (ho, wo), wt, counts, taps, None, (H, W))
                    gx = gx.reshape(B, 1, H, W) * ops.mask_to_f32(m8).reshape(B, 1, H, W)
            else:
                ops.wgrad_igemm(xin, gz, pk.fplan, pk.blks(cin, dev), pk.perm(dev), dw, x3=mode == "tf32x3")
                if ctx.needs_input_grad[0]:
                    if s == 1:
                        dx, _ = ops.conv_igemm(gz, pk.w_dgrad(weight, mode), pk.dplan, (H, W), code=m8.reshape(B, 1, H, W),
                                               lut=[0.0, 1.0])
                        gx = dx[:, 0].permute(0, 3, 1, 2).float()
                    else:
                        dx, _ = ops.conv_igemm(gz, pk.w_dgrad(weight, mode), pk.dplan, (H // 2, W // 2), code=m_split,
                                               lut=[0.0, 1.0])
                        gx = P.from_parity_split(dx).permute(0, 3, 1, 2).float()
        ctx.state = None
        if not module.batch_norm:
            dgam = dbet = None
        return gx, None, None, dw, dbias, dgam, dbet
