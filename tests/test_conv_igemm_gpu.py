"""GPU parity of the tcgen05 implicit-GEMM kernels (through the C ABI) against torch fp32 convs.

Inputs/weights are rounded to bf16 first so the only differences are fp32 accumulation order and the
final bf16 rounding of the output (tolerance 1e-2 relative to max|ref|, north-star bf16 bound).
"""
import pytest
import torch
import torch.nn.functional as F

from tg_b200 import ops, plan as P

pytestmark = pytest.mark.gpu


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


# (Cin, Cout, k, stride, pad, H, B)
FPROP = [
    (64, 64, 1, 1, 0, 16, 1),      # single tile, single K block: pure GEMM
    (64, 64, 3, 1, 1, 32, 2),      # dec1-like
    (192, 64, 3, 1, 1, 16, 2),     # dec2-like
    (128, 128, 3, 1, 1, 16, 3),    # vgg-like
    (1024, 512, 3, 1, 1, 8, 3),    # dec5..7-like, batch overhang (Bt=2, B=3)
    (64, 128, 5, 2, 2, 32, 2),     # enc2
    (128, 256, 5, 2, 2, 16, 2),    # enc3
    (256, 512, 3, 2, 1, 16, 2),    # enc4
    (512, 512, 3, 2, 1, 2, 5),     # enc7-like: 1x1 outputs, batch-tiled
    (64, 128, 4, 2, 1, 32, 2),     # D2
    (768, 256, 3, 1, 1, 8, 1),     # dec4
]


@pytest.mark.parametrize("Cin,Cout,k,s,p,H,B", FPROP)
def test_fprop(Cin, Cout, k, s, p, H, B):
    torch.manual_seed(0)
    dev = "cuda"
    x = torch.randn(B, Cin, H, H, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=dev)
    ref = F.conv2d(x.float(), w.float(), bias, s, p)
    Ho = ref.shape[2]
    pl = P.fprop_plan(k, s, p)
    xin = nhwc(x)
    xin = P.to_parity_split(xin) if s == 2 else xin.unsqueeze(1).contiguous()
    out, _ = ops.conv_igemm(xin, P.pack_w_fprop(w.float()), pl, (Ho, Ho), bias=bias)
    torch.cuda.synchronize()
    assert rel_err(out[:, 0], nhwc(ref)) < 1e-2


def test_fprop_epilogue_lut_affine_act_stats():
    torch.manual_seed(1)
    dev = "cuda"
    B, Cin, Cout, H, k = 2, 128, 256, 16, 3
    x = torch.randn(B, Cin, H, H, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * 9) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=dev)
    scale = torch.rand(Cout, device=dev) + 0.5
    shift = torch.randn(Cout, device=dev)
    code = torch.randint(0, 10, (B, 1, H, H), device=dev, dtype=torch.uint8)
    lut = P.ratio_lut(3)
    pl = P.fprop_plan(k, 1, 1)
    xin = nhwc(x).unsqueeze(1).contiguous()
    out, stats = ops.conv_igemm(xin, P.pack_w_fprop(w.float()), pl, (H, H), code=code, lut=lut, bias=bias,
                                scale=scale, shift=shift, act=2, slope=0.2, want_stats=True)
    z = nhwc(F.conv2d(x.float(), w.float(), bias, 1, 1)) * torch.tensor(lut, device=dev)[code.long()].reshape(B, H, H, 1)
    ref = F.leaky_relu(z * scale + shift, 0.2)
    assert rel_err(out[:, 0], ref) < 1e-2
    # BatchNorm partials: per-channel sum and sum of squares of the pre-affine value z
    s = stats.double().sum(0)
    assert rel_err(s[0], z.double().sum((0, 1, 2))) < 1e-3
    assert rel_err(s[1], (z.double() ** 2).sum((0, 1, 2))) < 1e-3


DGRAD = [
    (64, 64, 3, 1, 1, 32, 2),
    (192, 64, 3, 1, 1, 16, 2),
    (64, 128, 5, 2, 2, 32, 2),
    (256, 512, 3, 2, 1, 16, 2),
    (64, 128, 4, 2, 1, 32, 2),
    (512, 512, 3, 2, 1, 2, 5),
]


@pytest.mark.parametrize("Cin,Cout,k,s,p,H,B", DGRAD)
def test_dgrad(Cin, Cout, k, s, p, H, B):
    torch.manual_seed(2)
    dev = "cuda"
    x = torch.randn(B, Cin, H, H, device=dev, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cout * k * k) ** 0.5).bfloat16().float()
    y = F.conv2d(x, w, None, s, p)
    g = torch.randn_like(y).bfloat16()
    (dx,) = torch.autograd.grad(y, x, g.float())
    pl = P.dgrad_plan(k, s, p)
    wp = P.pack_w_dgrad(w, pl)
    mask = torch.randint(0, 2, (B, H, H), device=dev, dtype=torch.uint8)
    gin = nhwc(g).unsqueeze(1).contiguous()
    if s == 1:
        out, _ = ops.conv_igemm(gin, wp, pl, (H, H), code=mask.reshape(B, 1, H, H).contiguous(), lut=[0.0, 1.0])
        got = out[:, 0]
    else:
        msplit = P.to_parity_split(mask.unsqueeze(-1)).squeeze(-1).contiguous()
        out, _ = ops.conv_igemm(gin, wp, pl, (H // 2, H // 2), code=msplit, lut=[0.0, 1.0])
        got = P.from_parity_split(out)
    ref = nhwc(dx) * mask.unsqueeze(-1)
    assert rel_err(got, ref) < 1e-2


WGRAD = [
    (64, 64, 1, 1, 0, 8, 1),       # one K box, one M tile (num_blk = 1 -> padded pair)
    (64, 64, 3, 1, 1, 32, 2),      # dec1-like: odd block count (9)
    (192, 64, 3, 1, 1, 16, 2),
    (128, 256, 5, 2, 2, 16, 2),
    (256, 512, 3, 2, 1, 16, 2),
    (64, 128, 4, 2, 1, 32, 2),
    (512, 512, 3, 2, 1, 2, 5),
    (1024, 512, 3, 1, 1, 8, 3),
]


@pytest.mark.parametrize("Cin,Cout,k,s,p,H,B", WGRAD)
def test_wgrad(Cin, Cout, k, s, p, H, B):
    torch.manual_seed(3)
    dev = "cuda"
    x = torch.randn(B, Cin, H, H, device=dev).bfloat16()
    w = torch.randn(Cout, Cin, k, k, device=dev, requires_grad=True)
    y = F.conv2d(x.float(), w, None, s, p)
    g = torch.randn_like(y).bfloat16()
    (dw_ref,) = torch.autograd.grad(y, w, g.float())
    pl = P.fprop_plan(k, s, p)
    xin = nhwc(x)
    xin = P.to_parity_split(xin) if s == 2 else xin.unsqueeze(1).contiguous()
    blks = ops.wgrad_blk_table(pl, Cin, dev)
    perm = torch.tensor(pl.kpos, dtype=torch.int32, device=dev)
    dw = torch.full((Cout, Cin, k, k), 1.0, device=dev)
    ops.wgrad_igemm(xin, nhwc(g).unsqueeze(1).contiguous(), pl, blks, perm, dw, accumulate=False)
    assert rel_err(dw, dw_ref) < 1e-2
    ops.wgrad_igemm(xin, nhwc(g).unsqueeze(1).contiguous(), pl, blks, perm, dw, accumulate=True)
    assert rel_err(dw, 2 * dw_ref) < 1e-2


# (Cin, Cout, k, stride, pad, H, W, B, slope): data gradient with the activation-derivative gate of the tensor the
# gradient flows into, on every kernel that serves it: resident-weight halo (64->64), streaming halo (N = 128 with
# large weights, N = 64 with 192 input channels of the forward layer), stride-2 halo (four parity planes) and the
# generic kernel (N = 256). Non-square shapes give the halo kernels an odd number of 8x16 tiles.
GATED = [
    (64, 64, 3, 1, 1, 32, 32, 2, 0.0),
    (64, 64, 3, 1, 1, 16, 24, 1, 0.2),
    (128, 128, 3, 1, 1, 16, 24, 3, 0.0),
    (64, 192, 3, 1, 1, 16, 16, 2, 0.0),
    (64, 128, 4, 2, 1, 32, 32, 2, 0.2),
    (64, 128, 5, 2, 2, 32, 48, 1, 0.0),
    (256, 128, 3, 1, 1, 16, 16, 2, 0.0),
]


@pytest.mark.parametrize("Cin,Cout,k,s,p,H,W,B,slope", GATED)
def test_dgrad_gate(Cin, Cout, k, s, p, H, W, B, slope):
    torch.manual_seed(5)
    dev = "cuda"
    x = torch.randn(B, Cin, H, W, device=dev, requires_grad=True)
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cout * k * k) ** 0.5).bfloat16().float()
    y = F.conv2d(x, w, None, s, p)
    g = torch.randn_like(y).bfloat16()
    (dx,) = torch.autograd.grad(y, x, g.float())
    act_out = torch.randn(B, Cin, H, W, device=dev).bfloat16()          # output of the (Leaky)ReLU the input came through
    factor = torch.where(act_out.float() > 0, torch.ones((), device=dev), torch.full((), slope, device=dev))
    ref = nhwc(dx * factor)
    pl = P.dgrad_plan(k, s, p)
    wp = P.pack_w_dgrad(w, pl)
    gin = nhwc(g).unsqueeze(1).contiguous()
    if s == 1:
        gate = nhwc(act_out).unsqueeze(1).contiguous()
        out, _ = ops.conv_igemm(gin, wp, pl, (H, W), gate=gate, gate_slope=slope)
        got = out[:, 0]
    else:
        gate = P.to_parity_split(nhwc(act_out))
        out, _ = ops.conv_igemm(gin, wp, pl, (H // 2, W // 2), gate=gate, gate_slope=slope)
        got = P.from_parity_split(out)
    torch.cuda.synchronize()
    assert rel_err(got, ref) < 1e-2


# streaming-weights halo kernel with the full forward epilogue (ratio LUT, BN statistics) on a non-square grid
@pytest.mark.parametrize("Cin,Cout,H,W,B", [(192, 64, 16, 24, 3), (384, 128, 32, 16, 1)])
def test_fprop_stream_halo_lut_stats(Cin, Cout, H, W, B):
    torch.manual_seed(6)
    dev = "cuda"
    x = torch.randn(B, Cin, H, W, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=dev)
    code = torch.randint(0, 10, (B, 1, H, W), device=dev, dtype=torch.uint8)
    lut = P.ratio_lut(3)
    pl = P.fprop_plan(3, 1, 1)
    out, stats = ops.conv_igemm(nhwc(x).unsqueeze(1).contiguous(), P.pack_w_fprop(w.float()), pl, (H, W), code=code, lut=lut,
                                bias=bias, want_stats=True)
    z = nhwc(F.conv2d(x.float(), w.float(), bias, 1, 1)) * torch.tensor(lut, device=dev)[code.long()].reshape(B, H, W, 1)
    assert rel_err(out[:, 0], z) < 1e-2
    s = stats.double().sum(0)
    assert rel_err(s[0], z.double().sum((0, 1, 2))) < 1e-3
    assert rel_err(s[1], (z.double() ** 2).sum((0, 1, 2))) < 1e-3


@pytest.mark.parametrize("Cin,Cout,H,W,B", [(64, 64, 32, 16, 2), (128, 128, 32, 24, 3), (64, 128, 16, 8, 1), (64, 64, 24, 16, 2)])
def test_fused_maxpool_epilogue(Cin, Cout, H, W, B):
    """conv -> ReLU -> MaxPool2d(2,2) with the pool in the conv epilogue (VGG features[2..4], [7..9]): the pooled tensor is
    bit-identical to pooling the stored output, with and without writing the full-resolution tensor; a shape the halo
    kernels do not take (H % 16 != 0) goes through the stand-alone pooling kernel with the same result."""
    torch.manual_seed(5)
    dev = "cuda"
    x = torch.randn(B, Cin, H, W, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5).bfloat16()
    bias = torch.randn(Cout, device=dev)
    pl = P.fprop_plan(3, 1, 1)
    xin = nhwc(x).unsqueeze(1).contiguous()
    wp = P.pack_w_fprop(w.float())
    ref, _ = ops.conv_igemm(xin, wp, pl, (H, W), bias=bias, act=1)
    want = ops.maxpool2(ref[:, 0])
    full, _, pooled = ops.conv_igemm(xin, wp, pl, (H, W), bias=bias, act=1, pool="also")
    assert torch.equal(full, ref) and torch.equal(pooled[:, 0], want)
    none, _, pooled2 = ops.conv_igemm(xin, wp, pl, (H, W), bias=bias, act=1, pool="only")
    assert none is None and torch.equal(pooled2[:, 0], want)
    tref = F.max_pool2d(F.relu(F.conv2d(x.float(), w.float(), bias, 1, 1)), 2)
    assert rel_err(pooled[:, 0], nhwc(tref)) < 1e-2
