"""CPU emulation of the *index semantics* of the tcgen05 kernels (tests only).

These follow the formulas in include/terragan_b200.h literally (gather taps with zero fill, dot with
the packed weights) in fp32 on the CPU, so the tap tables / weight packing / parity-split layouts
of tg_b200.plan can be validated against torch.nn.functional without a GPU, and so the GPU tests
can compare the kernels with exactly the contraction they are specified to compute.
"""
import torch


def shift_gather(x, plane, dh, dw, Ho, Wo):
    """x: [B,P,H,W,C] -> [B,Ho,Wo,C] reading plane at (h+dh, w+dw), zero outside."""
    B, P, H, W, C = x.shape
    out = torch.zeros(B, Ho, Wo, C, dtype=x.dtype)
    h0, h1 = max(0, -dh), min(Ho, H - dh)
    w0, w1 = max(0, -dw), min(Wo, W - dw)
    if h1 > h0 and w1 > w0:
        out[:, h0:h1, w0:w1] = x[:, plane, h0 + dh:h1 + dh, w0 + dw:w1 + dw]
    return out


def conv_igemm(x, w_packed, plan, out_hw, code=None, lut=None, bias=None, scale=None, shift=None,
               act=0, slope=0.0):
    """fp32 emulation of tg_conv_igemm. Returns out [B,Po,Ho,Wo,N] (fp32, unrounded)."""
    x = x.float()
    w = w_packed.float()
    B, P, H, W, C = x.shape
    N = w.shape[0]
    Ho, Wo = out_hw
    out = torch.zeros(B, plan.out_planes, Ho, Wo, N)
    for (tb, tc, koff, op) in plan.subs:
        acc = torch.zeros(B, Ho, Wo, N)
        for t in range(tc):
            pl, dh, dw = plan.taps[tb + t]
            a = shift_gather(x, pl, dh, dw, Ho, Wo)
            wk = w[:, (koff + t) * C:(koff + t + 1) * C]
            acc += a @ wk.t()
        out[:, op] = acc
    if bias is not None:
        out = out + bias.float()
    if code is not None:
        out = out * torch.tensor(lut, dtype=torch.float32)[code.long()].reshape(B, plan.out_planes, Ho, Wo, 1)
    if scale is not None:
        out = out * scale.float()
    if shift is not None:
        out = out + shift.float()
    if act == 1:
        out = out.clamp_min(0)
    elif act == 2:
        out = torch.where(out > 0, out, out * slope)
    return out


def wgrad(x, g, plan):
    """fp32 emulation of tg_wgrad_igemm + reduce: returns dw [N, C, k*k] in kernel-position order."""
    x = x.float()
    g = g.float()
    B, P, H, W, C = x.shape
    _, _, Ho, Wo, N = g.shape
    T = len(plan.taps)
    dw = torch.zeros(N, C, T)
    gm = g.reshape(-1, N)
    for t, (pl, dh, dw_) in enumerate(plan.taps):
        a = shift_gather(x, pl, dh, dw_, Ho, Wo).reshape(-1, C)
        dw[:, :, plan.kpos[t]] = gm.t() @ a
    return dw
