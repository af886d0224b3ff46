"""Time one conv_igemm configuration in isolation (CUDA events). usage: time_conv.py Cin Cout k stride H B [reps]"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
from tg_b200 import ops, plan as P
Cin, Cout, k, s, H, B = [int(v) for v in sys.argv[1:7]]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
dev = "cuda"
pl = P.fprop_plan(k, s, k // 2)
shape = (B, 4, H // 2, H // 2, Cin) if s == 2 else (B, 1, H, H, Cin)
x = torch.randn(shape, device=dev).bfloat16()
w = torch.randn(Cout, Cin, k, k, device=dev)
wp = P.pack_w_fprop(w)
bias = torch.randn(Cout, device=dev)
Ho = H // s
out = torch.empty((B, 1, Ho, Ho, Cout), dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run():
    ops.conv_igemm(x, wp, pl, (Ho, Ho), bias=bias, act=1, out=out)
for _ in range(2): run()
ts = []
for _ in range(reps):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
fl = 2.0 * B * Ho * Ho * Cout * Cin * k * k
t = min(ts)
print(f"debug={os.environ.get('TG_CONV_DEBUG','0')} nohalo={os.environ.get('TG_NO_HALO','0')}: {t*1e3:.1f} us  {fl/t/1e9:.1f} TFLOP/s")
