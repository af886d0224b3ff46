"""Per-layer gradient diagnostic: CUDA generator backward vs the rounding-emulating oracle (smooth regime)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200")); sys.path.insert(0, ROOT)
from oracle import terra_oracle as O
from mvp_gan.src.models.generator import PConvUNet

DEV = "cuda"
H, B = int(sys.argv[1]) if len(sys.argv) > 1 else 256, 2
mode = (sys.argv[2] if len(sys.argv) > 2 else "train") == "train"
shift = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0


def smooth(sd):
    for k in sd:
        if k.endswith("bn.bias"):
            sd[k] = sd[k] + shift
    return sd


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-20)).item()


x, mask, target = O.make_tiles(70, B, H), O.make_mask(71, B, H, "rect"), O.make_tiles(72, B, H)
names = O._leaf_params(O.make_generator_state(1))
with O.rounding(O.bf16_ste):
    sd = O._require_grad(smooth(O.make_generator_state(1)))
    trace = {}
    o = O.pconv_unet(x * mask, mask, sd, mode, trace)
    for n, *_ in O.ENC + O.DEC:
        trace[n + ".z"].retain_grad()
        trace[n + ".y"].retain_grad()
    loss = ((o - target) ** 2).mean()
    loss.backward()
G = PConvUNet()
G.load_state_dict(smooth(O.make_generator_state(1)))
G.to(DEV).train(mode)
G._engine.debug = {}
G._trace = {}
out = G((x * mask).to(DEV), mask.to(DEV))
l2 = ((out - target.to(DEV)) ** 2).mean()
l2.backward()
print("out", rel(out, o), "loss", l2.item(), loss.item())
for n, cin, cout, k, s, p in O.ENC + O.DEC:
    # oracle: z.grad = dL/dz (z = ratio-scaled conv output); ours gz = dL/d(conv+bias) = dz * ratio
    msum = trace[n + ".msum"]
    ratio = (k * k / (msum + 1e-8)) * (msum > 0).float()
    gz_ref = (trace[n + ".z"].grad * ratio).permute(0, 2, 3, 1)
    gz = G._engine.debug[n + ".gz"][:, 0]
    y_err = rel(G._trace[n + ".y"].permute(0, 3, 1, 2), trace[n + ".y"])
    w = getattr(G, n).input_conv.weight.grad
    print(f"{n}: y {y_err:.4f}  gz {rel(gz, gz_ref):.4f}  dW {rel(w, sd[n + '.input_conv.weight'].grad):.4f}  "
          f"dgamma {rel(getattr(G, n).bn.weight.grad, sd[n + '.bn.weight'].grad):.4f}  "
          f"dbeta {rel(getattr(G, n).bn.bias.grad, sd[n + '.bn.bias'].grad):.4f}  |gz|max {gz_ref.abs().max():.3e}")

# ---- where are the largest gz errors? ----
for n in ("dec1", "dec4", "enc5"):
    k = 3
    msum = trace[n + ".msum"]
    ratio = (k * k / (msum + 1e-8)) * (msum > 0).float()
    gz_ref = (trace[n + ".z"].grad * ratio).permute(0, 2, 3, 1)
    gz = G._engine.debug[n + ".gz"][:, 0].float().cpu()
    err = (gz - gz_ref).abs()
    flat = err.flatten().topk(5).indices
    zref = trace[n + ".z"].detach().permute(0, 2, 3, 1)
    yref = trace[n + ".y"].detach().permute(0, 2, 3, 1)
    ygrad = trace[n + ".y"].grad.permute(0, 2, 3, 1)
    for f in flat.tolist():
        idx = list(torch.unravel_index(torch.tensor(f), err.shape))
        b, h, w, c = [int(i) for i in idx]
        print(n, (b, h, w, c), f"gz {gz[b,h,w,c]:.3e} ref {gz_ref[b,h,w,c]:.3e} s={int(msum[b,0,h,w])} z={zref[b,h,w,c]:.3f} "
              f"y={yref[b,h,w,c]:.3f} gy={ygrad[b,h,w,c]:.3e}")
