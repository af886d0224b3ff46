"""GPU parity of the bandwidth kernels (mask pyramid, BN fwd/bwd, resampling, 1-channel convs, fused
losses) through the C ABI against plain fp32 torch ops — the same ATen calls the reference makes."""
import pytest
import torch
import torch.nn.functional as F

from tg_b200 import ops, plan as P

pytestmark = pytest.mark.gpu
DEV = "cuda"


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


# ---------------- mask pyramid: bit-exact ----------------
@pytest.mark.parametrize("k,s,p", [(7, 2, 3), (5, 2, 2), (3, 2, 1), (3, 1, 1)])
def test_mask_window_sum_bit_exact(k, s, p):
    torch.manual_seed(0)
    B, H = 3, 64
    m = (torch.rand(B, 1, H, H, device=DEV) < 0.4).float()
    m[0, :, 10:50, 5:40] = 0
    ref = F.conv2d(m, torch.ones(1, 1, k, k, device=DEV), None, s, p)     # pconv.py:34 mask_conv
    mu = ops.mask_from_f32(m.reshape(B, H, H).contiguous())
    even = ref.shape[-1] % 2 == 0
    ssum, upd, upd_split, in_split = ops.mask_window_sum(mu, k, s, p, want_upd_split=even, want_in_split=True)
    assert torch.equal(ssum.float(), ref[:, 0])
    assert torch.equal(upd.float(), (ref[:, 0] > 0).float())
    if even:
        assert torch.equal(P.from_parity_split(upd_split.unsqueeze(-1)).squeeze(-1), upd)
    assert torch.equal(P.from_parity_split(in_split.unsqueeze(-1)).squeeze(-1), mu)
    assert torch.equal(ops.mask_to_f32(upd), (ref[:, 0] > 0).float())


# the packed-byte fast path (four outputs per thread; widths divisible by 4) and the generic kernel (everything else) on
# non-square grids, arbitrary non-zero mask bytes, a full mask (the largest counts) and a single hole at the corner
@pytest.mark.parametrize("k,s,p", [(7, 2, 3), (5, 2, 2), (3, 2, 1), (3, 1, 1)])
@pytest.mark.parametrize("B,H,W", [(2, 8, 24), (3, 40, 72), (1, 4, 4), (2, 6, 10), (2, 512, 512)])
@pytest.mark.parametrize("fill", ["random", "ones", "corner"])
def test_mask_window_sum_shapes(k, s, p, B, H, W, fill):
    g = torch.Generator(device="cpu").manual_seed(k * 100 + s * 10 + H)
    if fill == "random":
        mu = (torch.rand(B, H, W, generator=g) < 0.5).to(torch.uint8) * torch.randint(1, 256, (B, H, W), generator=g).to(torch.uint8)
    elif fill == "ones":
        mu = torch.full((B, H, W), 255, dtype=torch.uint8)
    else:
        mu = torch.ones(B, H, W, dtype=torch.uint8)
        mu[:, : min(H, 5), : min(W, 5)] = 0
        mu[:, -1, -1] = 0
    mu = mu.to(DEV)
    ref = F.conv2d((mu != 0).float().unsqueeze(1), torch.ones(1, 1, k, k, device=DEV), None, s, p)[:, 0]
    even = ref.shape[-1] % 2 == 0 and ref.shape[-2] % 2 == 0
    ssum, upd, upd_split, _ = ops.mask_window_sum(mu, k, s, p, want_upd_split=even)
    assert torch.equal(ssum.float(), ref)
    assert torch.equal(upd.float(), (ref > 0).float())
    if even:
        assert torch.equal(P.from_parity_split(upd_split.unsqueeze(-1)).squeeze(-1), upd)


def test_mask_merge_up_bit_exact():
    torch.manual_seed(1)
    B, H = 2, 32
    up = (torch.rand(B, 1, H // 2, H // 2, device=DEV) < 0.5).float()
    skip = (torch.rand(B, 1, H, H, device=DEV) < 0.5).float()
    ref = torch.max(F.interpolate(up, scale_factor=2, mode="nearest"), skip)   # generator.py:68,74
    got = ops.mask_merge_up(ops.mask_from_f32(up[:, 0].contiguous()), ops.mask_from_f32(skip[:, 0].contiguous()))
    assert torch.equal(got.float(), ref[:, 0])


# ---------------- BatchNorm forward / backward ----------------
# the last two shapes give every thread several pixels of a non-square grid (the kernels walk (b, h, w) incrementally with
# carries instead of dividing per pixel) and an odd number of pixels per channel-vector lane
@pytest.mark.parametrize("C,act,B,H,W", [(64, 1, 3, 16, 16), (128, 2, 3, 16, 16), (512, 1, 3, 16, 16), (64, 2, 6, 96, 80),
                                         (128, 1, 5, 112, 48)])
def test_bn_forward_backward(C, act, B, H, W):
    torch.manual_seed(2)
    z = torch.randn(B, C, H, W, device=DEV).bfloat16()
    gamma = (1 + 0.2 * torch.randn(C, device=DEV))
    beta = 0.2 * torch.randn(C, device=DEV)
    rm, rv = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    code = torch.randint(0, 10, (B, H, W), device=DEV, dtype=torch.uint8)
    lut = torch.tensor(P.ratio_lut(3), device=DEV)
    # stats partials as the conv epilogue would write them (2 rows)
    zf = nhwc(z.float())
    half = B * H * W // 2
    flat = zf.reshape(-1, C)
    partial = torch.stack([torch.stack([flat[:half].sum(0), (flat[:half] ** 2).sum(0)]),
                           torch.stack([flat[half:].sum(0), (flat[half:] ** 2).sum(0)])]).contiguous()
    scale, shift, mean, invstd = ops.bn_finalize(partial, B * H * W, gamma, beta, 1e-5, 0.1, rm, rv)
    y, ys = ops.bn_apply(nhwc(z), scale, shift, act, 0.2, code=code, want_nhwc=True, want_split=True, mask_split=True)
    # reference: the same chain in fp32 with autograd (z = pre-BN value incl. ratio; ratio enters via dz/dconv)
    conv = (z.float() / 1.0).clone().requires_grad_(True)   # treat z as ratio-scaled conv output
    rm2, rv2 = torch.zeros(C, device=DEV), torch.ones(C, device=DEV)
    g_ref, b_ref = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    yr = F.batch_norm(conv, rm2, rv2, g_ref, b_ref, True, 0.1, 1e-5)
    yr = F.relu(yr) if act == 1 else F.leaky_relu(yr, 0.2)
    assert rel_err(y, nhwc(yr)) < 1e-2
    m = (code > 0).float().unsqueeze(-1)
    assert rel_err(P.from_parity_split(ys), nhwc(yr) * m) < 1e-2
    assert rel_err(rm, rm2) < 1e-4 and rel_err(rv, rv2) < 1e-4
    # backward: two gradient sources (one plain, one parity-split), ratio folded into gz
    g0 = torch.randn(B, H, W, C, device=DEV).bfloat16()
    g1 = torch.randn(B, H, W, C, device=DEV).bfloat16()
    gtot = nchw(g0.float() + g1.float())
    (dz, dg, db) = torch.autograd.grad(yr, [conv, g_ref, b_ref], gtot)
    ratio = lut[code.long()].unsqueeze(-1)
    gz, dgamma, dbeta, dbias = ops.bn_bwd(ops.grad_src(g0), ops.grad_src(P.to_parity_split(g1), split=True), nhwc(z),
                                          scale, shift, mean, invstd, act, 0.2, code, lut)
    assert rel_err(gz[:, 0], nhwc(dz) * ratio) < 1.5e-2
    assert rel_err(dgamma, dg) < 1e-2 and rel_err(dbeta, db) < 1e-2
    assert rel_err(dbias, (nhwc(dz) * ratio).sum((0, 1, 2))) < 2e-2


def test_bn_eval_coeff():
    torch.manual_seed(3)
    C = 128
    g, b = torch.randn(C, device=DEV), torch.randn(C, device=DEV)
    rm, rv = torch.randn(C, device=DEV), torch.rand(C, device=DEV) + 0.1
    scale, shift = ops.bn_eval_coeff(g, b, rm, rv, 1e-5)
    x = torch.randn(2, C, 4, 4, device=DEV)
    ref = F.batch_norm(x, rm, rv, g, b, False, 0.1, 1e-5)
    got = x * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1)
    assert rel_err(got, ref) < 1e-5


# ---------------- decoder assembly / pooling ----------------
# the last case has more 8-channel items than threads in the grid on a non-square image with a channel-vector count that is
# not a power of two (the kernel walks (b, i, j, channel vector) with carries between a thread's items)
@pytest.mark.parametrize("Cu,Cs,B,h,w", [(64, 0, 2, 8, 8), (128, 64, 2, 8, 8), (512, 512, 2, 8, 8), (128, 64, 5, 72, 40)])
def test_upsample_concat_fwd_bwd(Cu, Cs, B, h, w):
    torch.manual_seed(4)
    up = torch.randn(B, Cu, h, w, device=DEV).bfloat16().float().requires_grad_(True)
    skip = torch.randn(B, Cs, 2 * h, 2 * w, device=DEV).bfloat16().float() if Cs else None
    mm = (torch.rand(B, 1, 2 * h, 2 * w, device=DEV) < 0.7).float()
    upf = F.interpolate(up, scale_factor=2, mode="bilinear", align_corners=False)      # generator.py:67
    merged = torch.cat([upf, skip], 1) if Cs else upf
    ref = merged * mm
    mm8 = ops.mask_from_f32(mm[:, 0].contiguous())
    got = ops.upsample_concat(nhwc(up.detach()).bfloat16(), nhwc(skip).bfloat16() if Cs else None, mm8)
    assert rel_err(got[:, 0], nhwc(ref)) < 1e-2
    g = torch.randn_like(ref).bfloat16().float() * mm      # dgrad epilogue already applied the mask
    (dup,) = torch.autograd.grad(ref, up, g)
    got_b = ops.upsample_concat_bwd(nhwc(g).bfloat16().unsqueeze(1).contiguous(), Cu)
    assert rel_err(got_b, nhwc(dup)) < 1e-2


def test_maxpool_fwd_bwd():
    torch.manual_seed(5)
    B, C, H = 2, 64, 16
    pre = torch.randn(B, C, H, H, device=DEV).bfloat16().float().requires_grad_(True)
    x = F.relu(pre)
    y = F.max_pool2d(x, 2, 2)
    g = torch.randn_like(y).bfloat16().float()
    (dpre,) = torch.autograd.grad(y, pre, g)
    xb = nhwc(x.detach()).bfloat16()
    assert torch.equal(ops.maxpool2(xb).float(), nhwc(y))
    got = ops.maxpool2_bwd(xb, nhwc(g).bfloat16(), relu_gate=True)
    assert rel_err(got, nhwc(dpre)) < 1e-6


# ---------------- 1-channel convolutions ----------------
@pytest.mark.parametrize("k,s,p,act,split", [(7, 2, 3, 0, False), (4, 2, 1, 2, True), (3, 1, 1, 1, False)])
def test_conv_c1_fwd_wgrad(k, s, p, act, split):
    torch.manual_seed(6)
    B, H = 2, 64
    x = torch.rand(B, 1, H, H, device=DEV)
    m = (torch.rand(B, 1, H, H, device=DEV) < 0.7).float()
    w = (torch.randn(64, 1, k, k, device=DEV) / k).requires_grad_(True)
    b = torch.randn(64, device=DEV).requires_grad_(True)
    msum = F.conv2d(m, torch.ones(1, 1, k, k, device=DEV), None, s, p)
    lut = torch.tensor(P.ratio_lut(k), device=DEV)
    code = msum[:, 0].to(torch.uint8).contiguous()
    z = F.conv2d(x * m, w, b, s, p) * lut[code.long()].unsqueeze(1)
    ref = z if act == 0 else (F.relu(z) if act == 1 else F.leaky_relu(z, 0.2))
    m8 = ops.mask_from_f32(m[:, 0].contiguous())
    out, stats = ops.conv_c1_fwd(x[:, 0].contiguous(), m8, k, s, p, w.detach().reshape(64, k * k).contiguous(),
                                 b.detach(), code=code, lut_dev=lut, act=act, slope=0.2, out_split=split,
                                 want_stats=True)
    got = P.from_parity_split(out) if split else out[:, 0]
    assert rel_err(got, nhwc(ref)) < 1e-2
    st = stats.double().sum(0)
    assert rel_err(st[0], z.double().sum((0, 2, 3))) < 1e-3
    assert rel_err(st[1], (z.double() ** 2).sum((0, 2, 3))) < 1e-3
    # weight / bias gradient for a given gz (bf16)
    gz = torch.randn_like(z).bfloat16()
    conv = F.conv2d(x * m, w, b, s, p)
    dw_ref, db_ref = torch.autograd.grad(conv, [w, b], gz.float())
    dw = torch.zeros(64, 1, k, k, device=DEV)
    db = torch.zeros(64, device=DEV)
    gin = nhwc(gz)
    gin = P.to_parity_split(gin) if split else gin
    ops.conv_c1_wgrad(x[:, 0].contiguous(), m8, k, s, p, gin.contiguous(), split, dw, db)
    assert rel_err(dw, dw_ref) < 1e-3 and rel_err(db, db_ref) < 1e-3


@pytest.mark.parametrize("B,H,W,mode", [(1, 16, 16, 0), (3, 48, 80, 1), (2, 30, 100, 0), (2, 128, 64, 1), (1, 8, 24, 0)])
def test_to1_3x3_fused_ragged_patches(B, H, W, mode):
    """64 -> 1, 3x3: the one-kernel path (14 x 14 interior patches) on sizes that are not multiples of the patch, non-square
    grids, both epilogues; H or W < 16 takes the two-kernel path (generator.py:56-62, losses.py:79-89 data gradient)."""
    torch.manual_seed(21)
    C = 64
    x = torch.randn(B, C, H, W, device=DEV).bfloat16().float()
    w = torch.randn(1, C, 3, 3, device=DEV) / 24
    b = torch.randn(1, device=DEV)
    mask = (torch.rand(B, 1, H, W, device=DEV) < 0.5).float()
    xin = torch.rand(B, 1, H, W, device=DEV) * mask
    conv = F.conv2d(x, w, b, 1, 1)
    ref = conv if mode == 0 else torch.sigmoid(conv) * (1 - mask) + xin * mask
    pl = P.fprop_plan(3, 1, 1)
    taps = [(dh, dw) for (_, dh, dw) in pl.taps]
    wt = w[0].permute(1, 2, 0).reshape(9, C).contiguous()
    m8 = ops.mask_from_f32(mask[:, 0].contiguous())
    out, sig = ops.conv_to1_fwd(nhwc(x).bfloat16(), False, (H, W), wt, [9], taps, b, (H, W), mode=mode,
                                mask=m8 if mode else None, xin=xin[:, 0].contiguous() if mode else None, want_sig=bool(mode))
    assert out.shape == (B, H, W)
    assert rel_err(out, ref[:, 0]) < 2e-3
    if mode:
        assert rel_err(sig, torch.sigmoid(conv)[:, 0]) < 2e-3


def test_final_conv_sigmoid_composite_and_backward():
    torch.manual_seed(7)
    B, H, C = 2, 32, 64
    d0 = torch.randn(B, C, H, H, device=DEV).bfloat16().float().requires_grad_(True)
    w = (torch.randn(1, C, 3, 3, device=DEV) / 24).requires_grad_(True)
    b = torch.randn(1, device=DEV).requires_grad_(True)
    mask = (torch.rand(B, 1, H, H, device=DEV) < 0.6).float()
    xin = torch.rand(B, 1, H, H, device=DEV) * mask
    out_ref = torch.sigmoid(F.conv2d(d0, w, b, 1, 1)) * (1 - mask) + xin * mask       # generator.py:56-62
    pl = P.fprop_plan(3, 1, 1)
    taps = [(dh, dw) for (_, dh, dw) in pl.taps]
    wt = w.detach()[0].permute(1, 2, 0).reshape(9, C).contiguous()
    m8 = ops.mask_from_f32(mask[:, 0].contiguous())
    out, sig = ops.conv_to1_fwd(nhwc(d0.detach()).bfloat16(), False, (H, H), wt, [9], taps, b.detach(), (H, H), mode=1,
                                mask=m8, xin=xin[:, 0].contiguous(), want_sig=True)
    assert rel_err(out, out_ref[:, 0]) < 2e-3
    g = torch.randn_like(out_ref)
    dd0, dw_ref, db_ref = torch.autograd.grad(out_ref, [d0, w, b], g)
    g_pre = ops.final_bwd_pre(g[:, 0].contiguous(), sig, m8)
    dx = ops.conv_to1_bwd_data(g_pre, wt, taps, (H, H), C)
    assert rel_err(dx, nhwc(dd0)) < 1e-2
    dw = torch.zeros(1, C, 3, 3, device=DEV)
    db = torch.zeros(1, device=DEV)
    ops.conv_to1_wgrad(nhwc(d0.detach()).bfloat16(), g_pre, taps, dw, db)
    assert rel_err(dw, dw_ref) < 2e-3 and rel_err(db, db_ref) < 2e-3


@pytest.mark.parametrize("B,H", [(2, 8), (3, 32), (1, 5), (7, 9)])
def test_d11_conv_and_backward(B, H):
    """Discriminator model[11] (512 -> 1, 4x4): the wide kernels of direct_conv.cu incl. ragged pixel counts."""
    torch.manual_seed(8)
    C = 512
    x = torch.randn(B, C, H, H, device=DEV).bfloat16().float().requires_grad_(True)
    w = (torch.randn(1, C, 4, 4, device=DEV) / 90).requires_grad_(True)
    b = torch.randn(1, device=DEV).requires_grad_(True)
    ref = F.conv2d(x, w, b, 1, 1)                                            # discriminator.py:22
    Ho = ref.shape[-1]
    pl = P.fprop_plan(4, 1, 1)
    taps = [(dh, dw) for (_, dh, dw) in pl.taps]
    wt = w.detach()[0].permute(1, 2, 0).reshape(16, C).contiguous()
    out, _ = ops.conv_to1_fwd(nhwc(x.detach()).bfloat16(), False, (H, H), wt, [16], taps, b.detach(), (Ho, Ho))
    assert rel_err(out, ref[:, 0]) < 2e-3
    g = torch.randn_like(ref)
    dx_ref, dw_ref, db_ref = torch.autograd.grad(ref, [x, w, b], g)
    dx = ops.conv_to1_bwd_data(g[:, 0].contiguous(), wt, taps, (H, H), C)
    assert rel_err(dx, nhwc(dx_ref)) < 1e-2
    dw = torch.zeros(1, C, 4, 4, device=DEV)
    db = torch.zeros(1, device=DEV)
    ops.conv_to1_wgrad(nhwc(x.detach()).bfloat16(), g[:, 0].contiguous(), taps, dw, db)
    assert rel_err(dw, dw_ref) < 2e-3 and rel_err(db, db_ref) < 2e-3


@pytest.mark.parametrize("k,s,p", [(4, 2, 1), (3, 1, 1)])
def test_c1_dgrad_via_to1(k, s, p):
    """data gradient of a 1->64 conv (Discriminator model[0], VGG conv0) = C->1 gather with tap classes."""
    torch.manual_seed(9)
    B, H = 2, 32
    x = torch.rand(B, 1, H, H, device=DEV, requires_grad=True)
    w = torch.randn(64, 1, k, k, device=DEV)
    y = F.conv2d(x, w, None, s, p)
    g = torch.randn_like(y).bfloat16()
    (dx_ref,) = torch.autograd.grad(y, x, g.float())
    pl = P.dgrad_plan(k, s, p)
    taps = [(dh, dw) for (_, dh, dw) in pl.taps]
    wt = w[:, 0].reshape(64, k * k)[:, pl.kpos].t().contiguous()           # [ntaps][64]
    counts = [c for (_, c, _, _) in pl.subs]
    Hg = y.shape[-1]
    gin = nhwc(g)
    for split in ([False, True] if s == 2 else [False]):
        gi = P.to_parity_split(gin) if split else gin
        out, _ = ops.conv_to1_fwd(gi.contiguous(), split, (Hg, Hg), wt, counts, taps, None, (H, H))
        assert rel_err(out, dx_ref[:, 0]) < 1e-3


# ---------------- fused losses ----------------
def _ref_terms(pred, target, mask):
    l1 = F.l1_loss(pred, target)
    x = pred * (1 - mask)
    B, _, H, W = x.shape
    tv = 2 * (((x[:, :, 1:] - x[:, :, :-1]) ** 2).sum() / x[:, :, 1:].numel() +
              ((x[:, :, :, 1:] - x[:, :, :, :-1]) ** 2).sum() / x[:, :, :, 1:].numel()) / B
    dil = F.max_pool2d(mask, 3, 1, 1)
    ero = 1 - F.max_pool2d(1 - mask, 3, 1, 1)
    bd = torch.clamp(dil - ero, 0, 1)
    bl = (torch.abs(pred - target) * bd).sum() / (bd.sum() + 1e-6)
    return l1, tv, bl, bd.sum()


def test_inpaint_loss_fwd_bwd():
    torch.manual_seed(10)
    B, H = 3, 64
    pred = torch.rand(B, 1, H, H, device=DEV, requires_grad=True)
    target = torch.rand(B, 1, H, H, device=DEV)
    mask = torch.ones(B, 1, H, H, device=DEV)
    mask[:, :, 10:40, 20:50] = 0
    mask[1, :, 0:5, :] = 0
    l1, tv, bl, nb = _ref_terms(pred, target, mask)
    terms = ops.inpaint_loss_fwd(pred.detach(), target, mask)
    for got, ref in zip(terms.tolist(), (l1, tv, bl, nb)):
        assert abs(got - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
    gt = torch.tensor([1.0, 0.1, 0.5], device=DEV)
    (gref,) = torch.autograd.grad(l1 * gt[0] + tv * gt[1] + bl * gt[2], pred)
    got = ops.inpaint_loss_bwd(pred.detach(), target, mask, terms, gt)
    assert rel_err(got, gref) < 1e-4
    # empty boundary -> zero boundary term and zero gradient from it (losses.py:411)
    ones = torch.ones_like(mask)
    t2 = ops.inpaint_loss_fwd(pred.detach(), target, ones)
    assert t2[2].item() == 0.0 and t2[3].item() == 0.0
    # human-region L1 (losses.py:172): flags = 1 | 2
    hm = 1 - mask
    t3 = ops.inpaint_loss_fwd(pred.detach(), target, hm, flags=3)
    assert abs(t3[0].item() - F.l1_loss(pred * hm, target * hm).item()) < 1e-6


def test_l1_bf16_fwd_bwd():
    torch.manual_seed(11)
    a = F.relu(torch.randn(2, 16, 16, 256, device=DEV)).bfloat16()
    b = F.relu(torch.randn(2, 16, 16, 256, device=DEV)).bfloat16()
    out = ops.l1_bf16_fwd(a, b)
    ref = (a.float() - b.float()).abs().mean()
    assert abs(out.item() - ref.item()) < 1e-5
    go = torch.tensor([0.1], device=DEV)
    ga = ops.l1_bf16_bwd(a, b, go, relu_gate=True)
    gref = 0.1 * torch.sign(a.float() - b.float()) / a.numel() * (a.float() > 0)
    assert rel_err(ga, gref) < 1e-2
