"""Time the 64->64 3x3 data-gradient launch of conv_igemm in isolation, plain / with mask code / with ReLU gate."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
from tg_b200 import ops, plan as P
C, H, B = 64, 512, 64
dev = "cuda"
pl = P.dgrad_plan(3, 1, 1)
g = torch.randn(B, 1, H, H, C, device=dev).bfloat16()
w = torch.randn(C, C, 3, 3, device=dev)
wd = P.pack_w_dgrad(w, pl)
code = torch.randint(0, 2, (B, H, H), device=dev, dtype=torch.uint8)
gate = torch.randn(B, 1, H, H, C, device=dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(name, fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"{name:28s} {min(ts)*1e3:8.1f} us")
timeit("dgrad plain", lambda: ops.conv_igemm(g, wd, pl, (H, H)))
timeit("dgrad + mask code", lambda: ops.conv_igemm(g, wd, pl, (H, H), code=code, lut=[0.0, 1.0]))
timeit("dgrad + relu gate", lambda: ops.conv_igemm(g, wd, pl, (H, H), gate=gate, gate_slope=0.0))
