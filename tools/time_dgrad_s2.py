"""Time the stride-2 data gradient of a 64->128 4x4 conv (Discriminator model[2]) in isolation."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
from tg_b200 import ops, plan as P
k = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B, H, Cin, Cout = 64, 128, 64, 128
dev = "cuda"
pl = P.dgrad_plan(k, 2, k // 2 if k == 5 else 1)
g = torch.randn(B, 1, H, H, Cout, device=dev).bfloat16()
w = torch.randn(Cout, Cin, k, k, device=dev)
wd = P.pack_w_dgrad(w, pl)
gate = torch.randn(B, 4, H, H, Cin, device=dev).bfloat16()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(name, fn):
    for _ in range(2): fn()
    ts = []
    for _ in range(4):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    fl = 2.0 * B * H * H * Cout * Cin * k * k
    print(f"{name:28s} debug={os.environ.get('TG_CONV_DEBUG','0')} {min(ts)*1e3:8.1f} us {fl/min(ts)/1e9:8.1f} TFLOP/s")
timeit(f"s2 dgrad k={k} plain", lambda: ops.conv_igemm(g, wd, pl, (H, H)))
timeit(f"s2 dgrad k={k} + leaky gate", lambda: ops.conv_igemm(g, wd, pl, (H, H), gate=gate, gate_slope=0.2))
