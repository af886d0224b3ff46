"""CTA-pair (cluster of 2, tcgen05 cta_group::2) variant of the N = 256 implicit-GEMM convolution kernel.

The library picks it when a launch has at least one 256-pixel unit per pair of SMs, which the small parity shapes of
tests/test_conv_igemm_gpu.py never reach; here (a) the same parity tests are re-run in a subprocess with
TG_CONV_PAIR_MIN=1, which forces the pair kernel onto every N % 256 == 0 shape (odd tile counts fall back), and (b) one
production-sized launch is compared with the single-CTA kernel (TG_NO_CONV_PAIR=1 in a subprocess) bit for bit.
Reference op: nn.Conv2d forward / backward-data of PConv2d.input_conv (mvp_gan/src/models/pconv.py:30).
"""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_parity_suite_with_pair_kernel_forced():
    env = dict(os.environ, TG_CONV_PAIR_MIN="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_conv_igemm_gpu.py"), "-x", "-q",
                        "-m", "gpu", "-k", "fprop or dgrad", "-p", "no:cacheprovider"],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


_SCRIPT = r"""
import os, sys, torch
sys.path.insert(0, os.path.join({root!r}, "terra-gan_b200"))
from tg_b200 import ops, plan as P
torch.manual_seed(5)
B, H, W, Cin, Cout = 8, 64, 64, 256, 512
x = torch.randn(B, 1, H, W, Cin, device="cuda").bfloat16()
w = (torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05)
plan = P.fprop_plan(3, 1, 1)
bias = torch.randn(Cout, device="cuda")
y, stats = ops.conv_igemm(x, P.pack_w_fprop(w), plan, (H, W), bias=bias, want_stats=True)
torch.cuda.synchronize()
torch.save({{"y": y.cpu(), "sum": stats.float().sum(0).cpu() if stats is not None else None}}, sys.argv[1])
"""


def test_production_size_matches_single_cta_kernel(tmp_path):
    """8 x 64 x 64 pixels, 256 -> 512 channels, bias + BatchNorm partial sums: 256 pair units >= 74, so the default run takes
    the pair kernel; the accumulation order per output element is the same (taps x channel blocks), so y is bit-identical."""
    outs = []
    for tag, extra in (("pair", {}), ("single", {"TG_NO_CONV_PAIR": "1"})):
        f = tmp_path / f"{tag}.pt"
        r = subprocess.run([sys.executable, "-c", _SCRIPT.format(root=ROOT), str(f)], env=dict(os.environ, **extra),
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
        outs.append(torch.load(f))
    a, b = outs
    assert torch.equal(a["y"], b["y"])
    assert torch.isfinite(a["y"].float()).all() and a["y"].float().abs().max() > 0
    # the per-CTA partial rows are grouped differently; their totals agree to fp32 summation error
    torch.testing.assert_close(a["sum"], b["sum"], rtol=1e-4, atol=1e-2)
