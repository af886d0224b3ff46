"""Time the three Discriminator model[11] kernels at the bench shape (B=64, 32x32x512)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
import torch
from tg_b200 import ops, plan as P
dev = "cuda"
B, H, C = 64, 32, 512
x = torch.randn(B, H, H, C, device=dev).bfloat16()
wt = torch.randn(16, C, device=dev) / 90
b = torch.zeros(1, device=dev)
taps = [(dh, dw) for (_, dh, dw) in P.fprop_plan(4, 1, 1).taps]
g = torch.randn(B, H - 1, H - 1, device=dev)
dw, db = torch.zeros(1, C, 4, 4, device=dev), torch.zeros(1, device=dev)
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
print("fwd  %.1f us" % t(lambda: ops.conv_to1_fwd(x, False, (H, H), wt, [16], taps, b, (H - 1, H - 1))))
print("bwd  %.1f us" % t(lambda: ops.conv_to1_bwd_data(g, wt, taps, (H, H), C)))
print("wgrad %.1f us" % t(lambda: ops.conv_to1_wgrad(x, g, taps, dw, db)))
