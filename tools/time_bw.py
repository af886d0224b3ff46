"""Time the streaming kernels (BatchNorm, upsample-concat, max-pool backward, L1 backward) in isolation at the B=64 shapes of
the train step (CUDA events, L2 flushed). A/B switches: TG_WAVE_GRID=0 (8 x SMs grids), TG_BN_REDUCE_OCC=1."""
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
# dec2: 128 up-sampled + 64 skip channels at 256^2
up2 = torch.randn(B, 128, 128, 128, device=dev).bfloat16(); sk2 = torch.randn(B, 256, 256, 64, device=dev).bfloat16()
mm2 = torch.randint(0, 2, (B, 256, 256), device=dev, dtype=torch.uint8)
timeit("upsample_concat dec2", lambda: ops.upsample_concat(up2, sk2, mm2), (B * 128 * 128 * 128 + B * 256 * 256 * (64 + 192)) * 2)
# encoder-type backward: skip gradient + parity-split gradient of the next layer (mode 7), 64 ch at 256^2
Hh = 256
z7 = torch.randn(B, Hh, Hh, C, device=dev).bfloat16()
ga = torch.randn(B, Hh, Hh, C, device=dev).bfloat16(); gb = torch.randn(B, 4, Hh // 2, Hh // 2, C, device=dev).bfloat16()
code7 = torch.randint(0, 50, (B, Hh, Hh), device=dev, dtype=torch.uint8); lut7 = torch.tensor(P.ratio_lut(7), device=dev)
def bwd7():
    ops.bn_bwd(ops.grad_src(ga), ops.grad_src(gb, split=True), z7, scale, shift, mean, invstd, 1, 0.0, code7, lut7)
try:
    timeit("bn_bwd enc1-type (2 sources)", bwd7, B * Hh * Hh * C * 12)
except Exception as e:
    print("bn_bwd enc1-type skipped:", e)
y2 = torch.randn(B, 512, 512, 64, device=dev).bfloat16(); gy = torch.randn(B, 256, 256, 64, device=dev).bfloat16()
timeit("maxpool2_bwd 512^2x64", lambda: ops.maxpool2_bwd(y2, gy), M * C * 4.5)
go = torch.ones(1, device=dev)
timeit("l1_bf16_bwd 512^2x64", lambda: ops.l1_bf16_bwd(y2, z, go), M * C * 6)
