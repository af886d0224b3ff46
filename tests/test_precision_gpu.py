"""The fp32-storage verification path (BASELINE.json north star: "TF32 path <= 1e-3").

tg_b200.precision selects it: fp32 activations, kind::tf32 tcgen05 MMAs ("tf32": one pass; "tf32x3": two-term
TF32 splits, fp32-grade products). The SAME kernel templates as the bf16 product path run with the storage type
swapped, driven by the SAME host scheduling (tg_b200.layers forward / backward), so these tests pin the
forward and — above all — the backward formulation of every layer against the fp32 oracle to a sharp bound:

  * kernel level: tf32 / tf32x3 implicit-GEMM fprop, dgrad and wgrad and the fp32 twins of the bandwidth kernels
    against torch fp32;
  * module level (the reference's own shapes): generator output, every loss term and EVERY parameter gradient of
    the adversarial step and of the human-guided step within 1e-3 (max-norm, per tensor) of the fp32 oracle in
    tf32x3 mode — a dgrad / wgrad / BN-backward / upsample-backward that was off by a few per cent fails here.
"""
import pytest
import torch
import torch.nn.functional as F

from oracle import terra_oracle as O
from tg_b200 import ops, plan as P, precision as PR
from mvp_gan.src.models.pconv import PConv2d
from mvp_gan.src.models.generator import PConvUNet
from mvp_gan.src.models.discriminator import Discriminator
from mvp_gan.src.utils.losses import HumanGuidedLoss, InpaintingLoss

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL_TF32 = 1e-3      # the north-star bound of the TF32 path


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel_err(a, b, scale=0.0):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / max(b.abs().max().item(), scale, 1e-30)).item()


@pytest.fixture(autouse=True)
def _fp32_reference_and_reset():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    PR.set_precision("bf16")


def _w(mode, w, plan=None):
    split = ops.split_hi_lo if mode == "tf32x3" else None
    return P.pack_w_fprop_f32(w, split) if plan is None else P.pack_w_dgrad_f32(w, plan, split)


# ---------------------------------------------------------------------------------------------------------------
# kernel level
# ---------------------------------------------------------------------------------------------------------------
FPROP = [(64, 64, 3, 1, 1, 32, 2), (192, 64, 3, 1, 1, 16, 2), (1024, 512, 3, 1, 1, 8, 3), (64, 128, 5, 2, 2, 32, 2),
         (256, 512, 3, 2, 1, 16, 2), (512, 512, 3, 2, 1, 2, 5), (64, 128, 4, 2, 1, 32, 2), (32, 128, 3, 1, 1, 16, 1)]


@pytest.mark.parametrize("mode,tol", [("tf32", 2e-3), ("tf32x3", 1e-4)])
@pytest.mark.parametrize("Cin,Cout,k,s,p,H,B", FPROP)
def test_fprop_f32(mode, tol, Cin, Cout, k, s, p, H, B):
    torch.manual_seed(0)
    x = torch.randn(B, Cin, H, H, device=DEV)
    w = torch.randn(Cout, Cin, k, k, device=DEV) / (Cin * k * k) ** 0.5
    bias = torch.randn(Cout, device=DEV)
    ref = F.conv2d(x.double(), w.double(), bias.double(), s, p).float()
    Ho = ref.shape[2]
    pl = P.fprop_plan(k, s, p)
    xin = nhwc(x)
    xin = P.to_parity_split(xin) if s == 2 else xin.unsqueeze(1).contiguous()
    out, _ = ops.conv_igemm(xin, _w(mode, w), pl, (Ho, Ho), bias=bias)
    assert out.dtype == torch.float32
    err = rel_err(out[:, 0], nhwc(ref))
    print(mode, (Cin, Cout, k, s), err)
    assert err < tol


@pytest.mark.parametrize("mode,tol", [("tf32", 2e-3), ("tf32x3", 1e-4)])
def test_fprop_f32_epilogue(mode, tol):
    """ratio LUT, bias, affine, LeakyReLU, BN partial sums in the fp32 epilogue."""
    torch.manual_seed(1)
    B, Cin, Cout, H, k = 2, 128, 256, 16, 3
    x = torch.randn(B, Cin, H, H, device=DEV)
    w = torch.randn(Cout, Cin, k, k, device=DEV) / (Cin * 9) ** 0.5
    bias, scale, shift = torch.randn(Cout, device=DEV), torch.rand(Cout, device=DEV) + 0.5, torch.randn(Cout, device=DEV)
    code = torch.randint(0, 10, (B, 1, H, H), device=DEV, dtype=torch.uint8)
    lut = P.ratio_lut(3)
    pl = P.fprop_plan(k, 1, 1)
    xin = nhwc(x).unsqueeze(1).contiguous()
    out, stats = ops.conv_igemm(xin, _w(mode, w), pl, (H, H), code=code, lut=lut, bias=bias, scale=scale, shift=shift,
                                act=2, slope=0.2, want_stats=True)
    z = nhwc(F.conv2d(x.double(), w.double(), bias.double(), 1, 1)).float() * torch.tensor(lut, device=DEV)[code.long()].reshape(B, H, H, 1)
    ref = F.leaky_relu(z * scale + shift, 0.2)
    assert rel_err(out[:, 0], ref) < tol
    s = stats.sum(0)
    assert rel_err(s[0], z.sum((0, 1, 2))) < 10 * tol and rel_err(s[1], (z * z).sum((0, 1, 2))) < tol


DGRAD = [(64, 64, 3, 1, 1, 32, 2), (128, 64, 5, 2, 2, 32, 2), (512, 256, 3, 2, 1, 16, 2), (128, 64, 4, 2, 1, 32, 2)]


@pytest.mark.parametrize("mode,tol", [("tf32", 2e-3), ("tf32x3", 1e-4)])
@pytest.mark.parametrize("Cout,Cin,k,s,p,H,B", DGRAD)
def test_dgrad_f32(mode, tol, Cout, Cin, k, s, p, H, B):
    torch.manual_seed(2)
    Ho = (H + 2 * p - k) // s + 1
    g = torch.randn(B, Cout, Ho, Ho, device=DEV)
    w = torch.randn(Cout, Cin, k, k, device=DEV) / (Cout * k * k) ** 0.5
    ref = F.conv_transpose2d(g.double(), w.double(), None, s, p, output_padding=H - ((Ho - 1) * s - 2 * p + k)).float()
    dpl = P.dgrad_plan(k, s, p)
    gate = torch.randn(B, H, H, Cin, device=DEV)
    gin = nhwc(g).unsqueeze(1).contiguous()
    hw = (H, H) if s == 1 else (H // 2, H // 2)
    gate_l = gate.unsqueeze(1).contiguous() if s == 1 else P.to_parity_split(gate)
    dx, _ = ops.conv_igemm(gin, _w(mode, w, dpl), dpl, hw, gate=gate_l, gate_slope=0.2)
    got = dx[:, 0] if s == 1 else P.from_parity_split(dx)
    want = nhwc(ref) * torch.where(gate > 0, 1.0, 0.2)
    err = rel_err(got, want)
    print(mode, (Cout, Cin, k, s), err)
    assert err < tol


WGRAD = [(64, 64, 3, 1, 1, 32, 2), (192, 64, 3, 1, 1, 16, 2), (64, 128, 5, 2, 2, 32, 2), (512, 512, 3, 2, 1, 4, 5),
         (1024, 512, 3, 1, 1, 8, 3), (128, 256, 4, 2, 1, 16, 2)]


@pytest.mark.parametrize("mode,tol", [("tf32", 2e-3), ("tf32x3", 1e-4)])
@pytest.mark.parametrize("Cin,Cout,k,s,p,H,B", WGRAD)
def test_wgrad_f32(mode, tol, Cin, Cout, k, s, p, H, B):
    torch.manual_seed(3)
    x = torch.randn(B, Cin, H, H, device=DEV)
    Ho = (H + 2 * p - k) // s + 1
    g = torch.randn(B, Cout, Ho, Ho, device=DEV)
    ref = torch.nn.grad.conv2d_weight(x.double(), (Cout, Cin, k, k), g.double(), s, p).float()
    pl = P.fprop_plan(k, s, p)
    xin = nhwc(x)
    xin = P.to_parity_split(xin) if s == 2 else xin.unsqueeze(1).contiguous()
    dw = torch.empty(Cout, Cin, k, k, device=DEV)
    perm = torch.tensor(pl.kpos, dtype=torch.int32, device=DEV)
    ops.wgrad_igemm(xin, nhwc(g).unsqueeze(1).contiguous(), pl, None, perm, dw, x3=mode == "tf32x3")
    err = rel_err(dw, ref)
    print(mode, (Cin, Cout, k, s), err)
    assert err < tol


def test_bandwidth_kernels_f32_match_their_bf16_twins_and_torch():
    """BN apply / backward, upsample-concat fwd/bwd, max-pool fwd/bwd, L1: the fp32 twins against torch fp32 (tight)
    and the bf16 product kernels against the fp32 twins (bf16 rounding of inputs and outputs only)."""
    torch.manual_seed(4)
    B, H, W, Cc = 2, 16, 32, 64
    z = torch.randn(B, H, W, Cc, device=DEV)
    scale, shift = torch.rand(Cc, device=DEV) + 0.5, torch.randn(Cc, device=DEV) * 0.3
    code = torch.randint(0, 3, (B, H, W), device=DEV, dtype=torch.uint8)
    y32, ys32 = ops.bn_apply(z, scale, shift, 1, 0.0, code=code, want_split=True, mask_split=True)
    ref = F.relu(z * scale + shift)
    assert rel_err(y32, ref) < 1e-6
    assert rel_err(P.from_parity_split(ys32), ref * (code > 0).unsqueeze(-1)) < 1e-6
    zb = z.bfloat16()
    y16, _ = ops.bn_apply(zb, scale, shift, 1, 0.0)
    assert rel_err(y16, F.relu(zb.float() * scale + shift)) < 5e-3
    # BN backward (train-mode statistics, ReLU, ratio LUT)
    lut = torch.tensor(P.ratio_lut(3), device=DEV)[:3].contiguous()
    zz = z.clone().requires_grad_(True)
    gam, bet = (torch.rand(Cc, device=DEV) + 0.5).requires_grad_(True), (torch.randn(Cc, device=DEV) * 0.3).requires_grad_(True)
    bias = torch.zeros(Cc, device=DEV, requires_grad=True)
    r = lut[code.long()].unsqueeze(-1)
    zr = (zz + bias) * r
    mean, var = zr.mean((0, 1, 2)), zr.var((0, 1, 2), unbiased=False)
    invstd = torch.rsqrt(var + 1e-5)
    yref = F.relu((zr - mean) * invstd * gam + bet)
    gy = torch.randn_like(yref)
    gz_ref, dgam_ref, dbet_ref, dbias_ref = torch.autograd.grad(yref, [zz, gam, bet, bias], gy)
    sc = (gam * invstd).detach()
    sh = (bet - mean * gam * invstd).detach()
    gz, dgam, dbet, dbias = ops.bn_bwd(ops.grad_src(gy.contiguous()), None, zr.detach().contiguous(), sc, sh, mean.detach(),
                                       invstd.detach(), 1, 0.0, code, lut)
    assert rel_err(gz[:, 0], gz_ref) < 2e-5 and rel_err(dgam, dgam_ref) < 2e-5 and rel_err(dbet, dbet_ref) < 2e-5
    assert rel_err(dbias, dbias_ref, dbet_ref.abs().max().item()) < 2e-5
    # upsample-concat forward / backward
    up, skip = torch.randn(B, 8, 8, 64, device=DEV), torch.randn(B, 16, 16, 64, device=DEV)
    mm = torch.randint(0, 2, (B, 16, 16), device=DEV, dtype=torch.uint8)
    m32 = ops.upsample_concat(up, skip, mm)
    upr = F.interpolate(up.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    ref = torch.cat([upr, skip], -1) * mm.unsqueeze(-1)
    assert rel_err(m32[:, 0], ref) < 1e-6
    assert rel_err(ops.upsample_concat(up.bfloat16(), skip.bfloat16(), mm)[:, 0], ref) < 8e-3
    d = torch.randn(B, 1, 16, 16, 128, device=DEV)
    upg = up.clone().requires_grad_(True)
    (F.interpolate(upg.permute(0, 3, 1, 2), scale_factor=2, mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
     * d[:, 0, :, :, :64]).sum().backward()
    assert rel_err(ops.upsample_concat_bwd(d, 64), upg.grad) < 1e-6
    # max-pool
    xp = torch.randn(B, 16, 16, 64, device=DEV)
    yp = ops.maxpool2(xp)
    assert torch.equal(yp, F.max_pool2d(xp.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1))
    gyp = torch.randn_like(yp)
    xg = xp.clone().requires_grad_(True)
    (F.max_pool2d(F.relu(xg).permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1) * gyp).sum().backward()
    assert rel_err(ops.maxpool2_bwd(xp, gyp, relu_gate=True), xg.grad) < 1e-6
    # L1 on features
    a, b = torch.randn(B, 8, 8, 256, device=DEV), torch.randn(B, 8, 8, 256, device=DEV)
    assert abs(ops.l1_bf16_fwd(a, b).item() - (a - b).abs().mean().item()) < 1e-6
    ga = ops.l1_bf16_bwd(a, b, torch.ones(1, device=DEV), relu_gate=True)
    assert rel_err(ga, torch.sign(a - b) / a.numel() * (a > 0)) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# module level: <= 1e-3 on outputs, losses and every gradient
# ---------------------------------------------------------------------------------------------------------------
PCONV_CASES = [(1, 64, 7, 2, 3, 1, 64, "iid"), (64, 128, 5, 2, 2, 2, 16, "rect"), (256, 512, 3, 2, 1, 2, 8, "large"),
               (192, 64, 3, 1, 1, 2, 16, "rect"), (64, 64, 3, 1, 1, 1, 32, "large")]


@pytest.mark.parametrize("case", PCONV_CASES)
@pytest.mark.parametrize("mode", ["tf32", "tf32x3"])
def test_pconv2d_layer_tf32(case, mode):
    """BASELINE.json config 1 and the other window shapes: one PConv2d layer fwd + bwd (train-mode BN)."""
    cin, cout, k, s, p, B, H, kind = case
    sd = O.make_pconv_state(100, cin, cout, k)
    x = torch.randn((B, cin, H, H), generator=torch.Generator().manual_seed(200))
    mask = O.make_mask(300, B, H, kind)
    names = ["input_conv.weight", "input_conv.bias", "bn.weight", "bn.bias"]
    osd = {k_: v.clone() for k_, v in sd.items()}
    for n in names:
        osd[n] = osd[n].requires_grad_(True)
    xr = x.clone().requires_grad_(True)
    y_ref, m_ref = O.pconv2d(xr, mask, osd, "", s, p, True)
    gy = torch.randn(y_ref.shape, generator=torch.Generator().manual_seed(400))
    g_ref = torch.autograd.grad(y_ref, [xr] + [osd[n] for n in names], gy)
    layer = PConv2d(cin, cout, k, s, p)
    layer.load_state_dict(sd)
    layer.to(DEV).train()
    xc = x.to(DEV).requires_grad_(True)
    with PR.precision(mode):
        y, m = layer(xc, mask.to(DEV))
        y.backward(gy.to(DEV))
    assert torch.equal(m.cpu(), m_ref)
    assert rel_err(y, y_ref) < (1e-4 if mode == "tf32x3" else 5e-3)
    got = [xc.grad, layer.input_conv.weight.grad, layer.input_conv.bias.grad, layer.bn.weight.grad, layer.bn.bias.grad]
    rows = []
    for gname, a_, r32 in zip(["dx", "dw", "db", "dgamma", "dbeta"], got, g_ref):
        scale = g_ref[4].abs().max().item() if gname == "db" else 0.0     # conv bias before train-mode BN: ~0
        rows.append((gname, rel_err(a_, r32, scale)))
    print(case[:5], mode, [(n, f"{e:.1e}") for n, e in rows])
    # one TF32 pass: ~5e-4 forward noise flips ~4e-4 of the ReLU gates, and with a RANDOM upstream gradient that is a
    # ~sqrt(4e-4) = 2 % perturbation of dw (see oracle.gate_tape); the sharp gradient bound belongs to tf32x3
    tol = 1e-4 if mode == "tf32x3" else 0.25
    assert all(e < tol for _, e in rows), rows


def _bn_companion(k):
    if k.endswith("input_conv.bias"):
        return k.replace("input_conv.bias", "bn.bias")
    for ci, bi in ((2, 3), (5, 6), (8, 9)):
        if k == f"model.{ci}.bias":
            return f"model.{bi}.bias"
    return None


def _grad_table(named_params, ref):
    rows = []
    for k, p in named_params:
        if k not in ref:
            continue
        assert p.grad is not None, k
        comp = _bn_companion(k)
        scale = ref[comp].abs().max().item() if comp is not None and comp in ref else 0.0
        rows.append((k, rel_err(p.grad, ref[k], scale)))
    rows.sort(key=lambda r: -r[1])
    return rows


def _modules():
    G = PConvUNet()
    G.load_state_dict(O.make_generator_state(1))
    D = Discriminator()
    D.load_state_dict(O.make_discriminator_state(2))
    return G.to(DEV).train(), D.to(DEV).train(), O.make_vgg_state(3)


def _run_adversarial(G, D, criterion, real_c, masks_c, lr=2e-4):
    import gates as GT
    return GT.run_adversarial(G, D, criterion, real_c, masks_c, lr)


def _table(got: dict, ref: dict):
    rows = []
    for k, g in got.items():
        if k not in ref:
            continue
        comp = _bn_companion(k)
        scale = ref[comp].abs().max().item() if comp is not None and comp in ref else 0.0
        rows.append((k, rel_err(g, ref[k], scale)))
    rows.sort(key=lambda r: -r[1])
    return rows


@pytest.mark.parametrize("H,B,kind", [(256, 2, "rect"), (512, 2, "large")])
def test_adversarial_step_tf32x3_every_gradient(H, B, kind):
    """train.py:179-225 at 256^2 and at the real 512^2 tile size in tf32x3 mode. Generator output, every loss term
    and the BN running statistics within 1e-3 (measured ~5e-5) of the plain fp32 oracle; EVERY generator /
    discriminator parameter gradient within 1e-3 of the fp32 oracle evaluated on the same branch decisions (gate
    tape: the decisions differ on ~1e-5 of the elements, all within 1e-3 sigma of the threshold); parameters after
    Adam. Without the tape the same comparison is printed: ~1e-2, the sensitivity of a piecewise-linear network to
    which side of zero a 1e-5-sized pre-activation falls."""
    import gates as GT
    real, masks = O.make_tiles(30, B, H), O.make_mask(31, B, H, kind)
    vgg = O.make_vgg_state(3)
    G, D, _ = _modules()
    criterion = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=torch.device(DEV), vgg_state_dict=vgg)
    GT.arm(G, D, criterion)
    with PR.precision("tf32x3"):
        got = _run_adversarial(G, D, criterion, real.to(DEV), masks.to(DEV))
    gates = GT.collect(G, D, criterion)
    gates["sign.pixel"] = torch.sign(got["gen"].cpu() - real)
    GT.disarm(G, D, criterion)
    g_sd, d_sd = O.make_generator_state(1), O.make_discriminator_state(2)
    with O.gate_tape(gates) as tape:
        r = O.adversarial_step(real, masks, g_sd, d_sd, vgg, lr=2e-4, opt_state={})
    r_plain = O.adversarial_step(real, masks, O.make_generator_state(1), O.make_discriminator_state(2), vgg)
    rows_g, rows_d = _table(got["g_grads"], r["g_grads"]), _table(got["d_grads"], r["d_grads"])
    plain_g = _table(got["g_grads"], r_plain["g_grads"])
    print(f"{H}x{H} tf32x3: gen {rel_err(got['gen'], r_plain['gen']):.1e};", GT.summarize(tape))
    print("  same branch decisions: worst G grads", [(n, f"{e:.1e}") for n, e in rows_g[:4]],
          "worst D grads", [(n, f"{e:.1e}") for n, e in rows_d[:3]])
    print("  plain oracle (own decisions): worst G grads", [(n, f"{e:.1e}") for n, e in plain_g[:4]])
    assert rel_err(got["gen"], r_plain["gen"]) < TOL_TF32
    for key in ("g_loss", "g_adv", "d_loss"):
        assert abs(got[key].item() - r_plain[key].item()) < TOL_TF32 * abs(r_plain[key].item()), key
    flips = sum(v[0] for v in tape.report.values())
    total = sum(v[1] for v in tape.report.values())
    assert flips < 1e-4 * total, GT.summarize(tape)
    assert max(v[2] for v in tape.report.values()) < 1e-2, GT.summarize(tape)       # all flipped elements sit at the threshold
    assert len(rows_g) == 58 and len(rows_d) == 16
    assert rows_g[0][1] < TOL_TF32, rows_g[:8]
    assert rows_d[0][1] < TOL_TF32, rows_d[:8]
    assert plain_g[0][1] < 0.3, plain_g[:8]
    for sd_ref, mod in ((g_sd, G), (d_sd, D)):
        for k, v in mod.state_dict().items():
            if "running_" in k:
                assert rel_err(v, sd_ref[k]) < TOL_TF32, k
    # parameters after one Adam step: the first step moves every weight by lr * g / (|g| + eps), i.e. by ~lr * sign(g):
    # compare where the reference gradient is clearly non-zero (an element at the 1e-3 error level may change sign)
    worst = 0.0
    for k, p in G.named_parameters():
        if k in r["g_grads"] and p.requires_grad:
            gr = r["g_grads"][k].abs()
            big = (gr > 0.05 * gr.max()) & (gr > 1e-6)
            if big.any():
                worst = max(worst, ((p.detach().cpu() - g_sd[k]).abs()[big].max() / 2e-4).item())
    print("  Adam: worst parameter deviation in units of lr:", worst)
    assert worst < 1e-2


def test_human_guided_step_tf32x3_every_gradient():
    import gates as GT
    H, B = 256, 2
    images, masks = O.make_tiles(40, B, H), O.make_mask(41, B, H, "large")
    human = 1 - O.make_mask(42, B, H, "rect")
    vgg = O.make_vgg_state(3)
    G, _, _ = _modules()
    config = {"training": {"loss_weights": {"perceptual": 0.1, "tv": 0.1, "boundary": 0.5},
                           "modes": {"human_guided": {"human_feedback_weight": 0.3, "base_loss_weight": 0.7,
                                                      "learning_rate": 1e-4}}}}
    criterion = HumanGuidedLoss(config, device=torch.device(DEV), vgg_state_dict=vgg)
    ic, mc, hc = images.to(DEV), masks.to(DEV), human.to(DEV)
    vcrit = criterion                 # HumanGuidedLoss subclasses InpaintingLoss: one VGG engine
    GT.arm(G, None, vcrit)
    with PR.precision("tf32x3"):
        generated = G(ic * mc, mc)
        loss = criterion(generated, ic, mc, {"mask": hc})
        loss.backward()
    gates = GT.collect(G, None, vcrit)
    gates["sign.pixel"] = torch.sign(generated.detach().cpu() - images)
    GT.disarm(G, None, vcrit)
    with O.gate_tape(gates) as tape:
        r = O.human_guided_step(images, masks, human, O.make_generator_state(1), vgg, lr=1e-4, opt_state={})
    rows = _table({k: p.grad for k, p in G.named_parameters() if p.grad is not None}, r["g_grads"])
    print("H-G tf32x3: gen", rel_err(generated, r["gen"]), GT.summarize(tape), "worst grads",
          [(n, f"{e:.1e}") for n, e in rows[:5]])
    assert rel_err(generated, r["gen"]) < TOL_TF32
    assert abs(loss.item() - r["loss"].item()) < TOL_TF32 * abs(r["loss"].item())
    assert len(rows) == 58 and rows[0][1] < TOL_TF32, rows[:8]


def test_single_pass_tf32_outputs_and_losses():
    """One kind::tf32 pass per product (what "TF32" means in cuDNN terms): generator output and loss terms within
    the north star's 1e-3; gradients within 1e-2 on the same branch decisions."""
    import gates as GT
    H, B = 256, 2
    real, masks = O.make_tiles(30, B, H), O.make_mask(31, B, H, "rect")
    vgg = O.make_vgg_state(3)
    G, D, _ = _modules()
    criterion = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=torch.device(DEV), vgg_state_dict=vgg)
    GT.arm(G, D, criterion)
    with PR.precision("tf32"):
        got = _run_adversarial(G, D, criterion, real.to(DEV), masks.to(DEV))
    gates = GT.collect(G, D, criterion)
    gates["sign.pixel"] = torch.sign(got["gen"].cpu() - real)
    GT.disarm(G, D, criterion)
    with O.gate_tape(gates) as tape:
        r = O.adversarial_step(real, masks, O.make_generator_state(1), O.make_discriminator_state(2), vgg)
    r_plain = O.adversarial_step(real, masks, O.make_generator_state(1), O.make_discriminator_state(2), vgg)
    rows = _table(got["g_grads"], r["g_grads"])
    print("tf32 (1 pass): gen", rel_err(got["gen"], r_plain["gen"]), GT.summarize(tape), "worst G grads",
          [(n, f"{e:.1e}") for n, e in rows[:5]])
    assert rel_err(got["gen"], r_plain["gen"]) < 2e-3
    for key in ("g_loss", "g_adv", "d_loss"):
        assert abs(got[key].item() - r_plain[key].item()) < TOL_TF32 * abs(r_plain[key].item()), key
    assert rows[0][1] < 1e-2, rows[:8]


def test_eval_inference_tf32x3_and_folded_epilogue():
    """evaluate.py:47-50 (eval, no_grad): fp32 path within 1e-4 of the oracle; exercises the folded BN+ReLU epilogue."""
    B, H = 2, 128
    x, mask = O.make_tiles(80, B, H), O.make_mask(81, B, H, "large")
    with torch.no_grad():
        ref = O.pconv_unet(x * mask, mask, O.make_generator_state(1), False)
    G = PConvUNet()
    G.load_state_dict(O.make_generator_state(1))
    G.to(DEV).eval()
    with torch.no_grad(), PR.precision("tf32x3"):
        out = G((x * mask).to(DEV), mask.to(DEV))
    assert rel_err(out, ref) < 1e-4
    with torch.no_grad():
        out16 = G((x * mask).to(DEV), mask.to(DEV))
    assert rel_err(out16, ref) < 1e-2
