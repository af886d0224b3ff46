"""Diagnostics for the tcgen05 kernels on a GPU box: runs the smallest GEMM-shaped cases and saves
inputs / outputs / references under gpurun_out/ so layout or descriptor errors can be analysed offline."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
from tg_b200 import ops, plan as P  # noqa: E402

out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
dev = "cuda"
torch.backends.cudnn.allow_tf32 = False
res = {}


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def rel(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-12)).item()


print("sms", ops.num_sms(), torch.cuda.get_device_name(0), flush=True)
# 1. pure GEMM through fprop: 128 pixels x 64 ch, N=64, 1 tap
for (Cin, Cout, k, H, W, B) in [(64, 64, 1, 8, 16, 1), (128, 64, 1, 8, 16, 1), (64, 256, 1, 8, 16, 1), (64, 64, 3, 8, 16, 1)]:
    torch.manual_seed(0)
    x = torch.randn(B, Cin, H, W, device=dev).bfloat16()
    w = (torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5).bfloat16()
    ref = nhwc(F.conv2d(x.float(), w.float(), None, 1, k // 2))
    pl = P.fprop_plan(k, 1, k // 2)
    out, _ = ops.conv_igemm(nhwc(x).unsqueeze(1).contiguous(), P.pack_w_fprop(w.float()), pl, (H, W))
    torch.cuda.synchronize()
    e = rel(out[:, 0], ref)
    print(f"fprop Cin={Cin} Cout={Cout} k={k}: rel_err={e:.3e}", flush=True)
    res[f"fprop_{Cin}_{Cout}_{k}"] = dict(x=x.cpu(), w=w.cpu(), out=out.cpu(), ref=ref.cpu())

# 2. pure GEMM through wgrad: K = 64 pixels
for (Cin, Cout, k, H, W, B) in [(64, 64, 1, 8, 8, 1), (128, 64, 1, 8, 8, 1), (64, 256, 1, 8, 8, 2), (64, 64, 3, 8, 8, 1)]:
    torch.manual_seed(0)
    x = torch.randn(B, Cin, H, W, device=dev).bfloat16()
    w = torch.randn(Cout, Cin, k, k, device=dev, requires_grad=True)
    y = F.conv2d(x.float(), w, None, 1, k // 2)
    g = torch.randn_like(y).bfloat16()
    (dw_ref,) = torch.autograd.grad(y, w, g.float())
    pl = P.fprop_plan(k, 1, k // 2)
    blks = ops.wgrad_blk_table(pl, Cin, dev)
    perm = torch.tensor(pl.kpos, dtype=torch.int32, device=dev)
    dw = torch.zeros(Cout, Cin, k, k, device=dev)
    ops.wgrad_igemm(nhwc(x).unsqueeze(1).contiguous(), nhwc(g).unsqueeze(1).contiguous(), pl, blks, perm, dw)
    torch.cuda.synchronize()
    e = rel(dw, dw_ref)
    print(f"wgrad Cin={Cin} Cout={Cout} k={k}: rel_err={e:.3e}", flush=True)
    res[f"wgrad_{Cin}_{Cout}_{k}"] = dict(x=x.cpu(), g=g.cpu(), dw=dw.cpu(), ref=dw_ref.detach().cpu())

torch.save(res, os.path.join(out_dir, "diag_igemm.pt"))
print("saved", flush=True)
