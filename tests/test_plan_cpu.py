"""Host planning logic (tap tables, parity split, weight packing) vs torch.nn.functional, on CPU."""
import pytest
import torch
import torch.nn.functional as F

from tg_b200 import plan as P
import emulate as E

CASES = [  # (k, stride, pad) of every conv on the path: pconv.py / generator.py:13-28, discriminator.py:11,22
    (3, 1, 1), (3, 2, 1), (5, 2, 2), (7, 2, 3), (4, 2, 1), (4, 1, 1),
]


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("k,s,p", CASES)
def test_fprop_plan_matches_conv2d(k, s, p):
    torch.manual_seed(0)
    B, Cin, Cout, H = 2, 3, 5, 12
    x = torch.randn(B, Cin, H, H)
    w = torch.randn(Cout, Cin, k, k)
    ref = F.conv2d(x, w, None, s, p)
    Ho = ref.shape[2]
    pl = P.fprop_plan(k, s, p)
    xin = nhwc(x)
    xin = P.to_parity_split(xin) if s == 2 else xin.unsqueeze(1)
    out = E.conv_igemm(xin, P.pack_w_fprop(w).float(), pl, (Ho, Ho))
    # packed weights are bf16: compare against conv with bf16-rounded weights
    ref = F.conv2d(x, w.bfloat16().float(), None, s, p)
    torch.testing.assert_close(out[:, 0], nhwc(ref), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("k,s,p", CASES)
def test_dgrad_plan_matches_autograd(k, s, p):
    torch.manual_seed(1)
    B, Cin, Cout, H = 2, 4, 3, 12
    x = torch.randn(B, Cin, H, H, requires_grad=True)
    w = torch.randn(Cout, Cin, k, k).bfloat16().float()
    y = F.conv2d(x, w, None, s, p)
    g = torch.randn_like(y)
    (dx,) = torch.autograd.grad(y, x, g)
    pl = P.dgrad_plan(k, s, p)
    wp = P.pack_w_dgrad(w, pl).float()
    Hg = y.shape[2]
    if s == 1:
        if Hg != H:   # k4 s1 p1 shrinks the image (D11): dgrad grid is the *input* grid
            out = E.conv_igemm(nhwc(g).unsqueeze(1), wp, pl, (H, H))
        else:
            out = E.conv_igemm(nhwc(g).unsqueeze(1), wp, pl, (H, H))
        got = out[:, 0]
    else:
        out = E.conv_igemm(nhwc(g).unsqueeze(1), wp, pl, (H // 2, H // 2))
        got = P.from_parity_split(out)
    torch.testing.assert_close(got, nhwc(dx), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("k,s,p", CASES)
def test_wgrad_plan_matches_autograd(k, s, p):
    torch.manual_seed(2)
    B, Cin, Cout, H = 2, 4, 3, 12
    x = torch.randn(B, Cin, H, H)
    w = torch.randn(Cout, Cin, k, k, requires_grad=True)
    y = F.conv2d(x, w, None, s, p)
    g = torch.randn_like(y)
    (dw,) = torch.autograd.grad(y, w, g)
    pl = P.fprop_plan(k, s, p)
    xin = nhwc(x)
    xin = P.to_parity_split(xin) if s == 2 else xin.unsqueeze(1)
    got = E.wgrad(xin, nhwc(g).unsqueeze(1), pl)
    torch.testing.assert_close(got.reshape(Cout, Cin, k, k), dw, rtol=1e-4, atol=1e-4)


def test_parity_split_roundtrip():
    x = torch.arange(2 * 6 * 8 * 3, dtype=torch.float32).reshape(2, 6, 8, 3)
    s = P.to_parity_split(x)
    assert s.shape == (2, 4, 3, 4, 3)
    assert torch.equal(s[:, 2 * 1 + 0, 1, 2], x[:, 3, 4])
    assert torch.equal(P.from_parity_split(s), x)


@pytest.mark.parametrize("k", [3, 5, 7])
def test_ratio_lut_matches_reference_expression(k):
    # pconv.py:38-40 evaluated on every possible window count
    s = torch.arange(0, k * k + 1, dtype=torch.float32)
    ref = (k * k) / (s + 1e-8) * (s > 0).float()
    lut = torch.tensor(P.ratio_lut(k))
    assert torch.equal(lut, ref)
    assert lut[0] == 0 and lut[k * k] == 1.0


def test_pack_scatter_index_is_the_packing_permutation():
    """tg_b200.optim.Adam writes packed[dst[i]] = bf16(w.flat[i]); dst must reproduce pack_w_fprop / pack_w_dgrad."""
    torch.manual_seed(0)
    for (co, ci, k, s, p) in [(8, 4, 3, 1, 1), (6, 2, 4, 2, 1), (4, 3, 5, 2, 2)]:
        w = torch.randn(co, ci, k, k)
        for plan in (None, P.dgrad_plan(k, s, p)):
            dst = P.pack_scatter_index(w.shape, plan, "cpu")
            want = (P.pack_w_fprop(w) if plan is None else P.pack_w_dgrad(w, plan)).reshape(-1)
            got = torch.empty_like(want)
            got[dst.long()] = w.reshape(-1).to(torch.bfloat16)
            assert sorted(dst.tolist()) == list(range(w.numel()))
            assert torch.equal(got, want)
