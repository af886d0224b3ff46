"""Time the bandwidth kernels in isolation at the B=64 / 512x512 shapes of the train step (CUDA events)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
from tg_b200 import ops, plan as P
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

def timeit(name, fn, nbytes, reps=4):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print(f"{name:34s} {t*1e3:9.1f} us   {nbytes/t/1e6:8.1f} GB/s (algorithmic)")

H = 512; C = 64
M = B * H * H
z = torch.randn(B, H, H, C, device=dev).bfloat16()
g = torch.randn(B, H, H, C, device=dev).bfloat16()
scale = torch.rand(C, device=dev) + 0.5; shift = torch.randn(C, device=dev)
mean = torch.randn(C, device=dev); invstd = torch.rand(C, device=dev) + 0.5
code = torch.randint(0, 10, (B, H, H), device=dev, dtype=torch.uint8)
lut = torch.tensor(P.ratio_lut(3), device=dev)
timeit("bn_apply dec1", lambda: ops.bn_apply(z, scale, shift, 1), M * C * 4)
gs = ops.grad_src(g)
def bwd():
    ops.bn_bwd(gs, None, z, scale, shift, mean, invstd, 1, 0.0, code, lut)
timeit("bn_bwd (reduce+fin+apply) dec1", bwd, M * C * 10)
# individual pieces
import ctypes as Cc
from tg_b200._lib import lib, ptr, stream_ptr, check
rows_cap = ops.num_sms() * 4
partial = torch.empty((rows_cap, 5, C), dtype=torch.float32, device=dev)
used = Cc.c_int(0)
def red():
    check(lib().tg_bn_bwd_reduce(Cc.byref(gs), None, ptr(z), B, H, H, C, ptr(scale), ptr(shift), 1, 0.0, ptr(code), ptr(lut), ptr(partial), rows_cap, Cc.byref(used), stream_ptr()), "r")
timeit("  bn_bwd_reduce dec1", red, M * C * 4)
outs = torch.zeros((8, C), dtype=torch.float32, device=dev); gz = torch.empty(B, 1, H, H, C, dtype=torch.bfloat16, device=dev)
def app():
    check(lib().tg_bn_bwd_apply(Cc.byref(gs), None, ptr(z), B, H, H, C, ptr(shift), ptr(outs), 1, 0.0, ptr(code), ptr(lut), ptr(gz), stream_ptr()), "a")
timeit("  bn_bwd_apply dec1", app, M * C * 6)
# upsample concat dec1 (64ch up only) and dec2 (128 + 64)
up = torch.randn(B, 256, 256, 64, device=dev).bfloat16()
mm = torch.randint(0, 2, (B, 512, 512), device=dev, dtype=torch.uint8)
timeit("upsample_concat dec1", lambda: ops.upsample_concat(up, None, mm), (B * 256 * 256 * 64 + M * C) * 2)
dm = torch.randn(B, 1, 512, 512, 64, device=dev).bfloat16()
timeit("upsample_concat_bwd dec1", lambda: ops.upsample_concat_bwd(dm, 64), (B * 256 * 256 * 64 + M * C) * 2)
# final conv trio
pl = P.fprop_plan(3, 1, 1); taps = [(a, b) for (_, a, b) in pl.taps]
wt = torch.randn(9, 64, device=dev); bias = torch.randn(1, device=dev)
m8 = torch.randint(0, 2, (B, H, H), device=dev, dtype=torch.uint8); xin = torch.rand(B, H, H, device=dev)
timeit("final conv fwd (3x3 C64->1)", lambda: ops.conv_to1_fwd(z, False, (H, H), wt, [9], taps, bias, (H, H), mode=1, mask=m8, xin=xin, want_sig=True), M * C * 2)
gpre = torch.randn(B, H, H, device=dev)
timeit("final conv bwd_data", lambda: ops.conv_to1_bwd_data(gpre, wt, taps, (H, H), 64), M * C * 2)
dw = torch.zeros(1, 64, 3, 3, device=dev); db = torch.zeros(1, device=dev)
timeit("final conv wgrad", lambda: ops.conv_to1_wgrad(z, gpre, taps, dw, db), M * C * 2)
# 1-channel convs
x = torch.rand(B, H, H, device=dev)
w3 = torch.randn(64, 9, device=dev); b64 = torch.randn(64, device=dev)
timeit("c1_fwd 3x3 s1 (VGG conv0)", lambda: ops.conv_c1_fwd(x, None, 3, 1, 1, w3, b64, act=1), M * C * 2)
w7 = torch.randn(64, 49, device=dev)
s7 = torch.randint(0, 50, (B, 256, 256), device=dev, dtype=torch.uint8); lut7 = torch.tensor(P.ratio_lut(7), device=dev)
timeit("c1_fwd 7x7 s2 (enc1)", lambda: ops.conv_c1_fwd(x, m8, 7, 2, 3, w7, b64, code=s7, lut_dev=lut7, want_stats=True), B * 256 * 256 * 64 * 2)
w4 = torch.randn(64, 16, device=dev)
timeit("c1_fwd 4x4 s2 (D0)", lambda: ops.conv_c1_fwd(x, None, 4, 2, 1, w4, b64, act=2, slope=0.2, out_split=True), B * 256 * 256 * 64 * 2)
g1 = torch.randn(B, 1, 256, 256, 64, device=dev).bfloat16()
dw7 = torch.zeros(64, 1, 7, 7, device=dev)
timeit("c1_wgrad 7x7 s2 (enc1)", lambda: ops.conv_c1_wgrad(x, m8, 7, 2, 3, g1, False, dw7, None), B * 256 * 256 * 64 * 2)
g4 = torch.randn(B, 4, 128, 128, 64, device=dev).bfloat16()
dw4 = torch.zeros(64, 1, 4, 4, device=dev); db4 = torch.zeros(64, device=dev)
timeit("c1_wgrad 4x4 s2 (D0)", lambda: ops.conv_c1_wgrad(x, None, 4, 2, 1, g4, True, dw4, db4), B * 256 * 256 * 64 * 2)
y2 = torch.randn(B, 512, 512, 64, device=dev).bfloat16()
timeit("maxpool2 512^2x64", lambda: ops.maxpool2(y2), M * C * 2.5)
