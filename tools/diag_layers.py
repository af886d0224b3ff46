"""Diagnostic: per-layer comparison of dL/d(conv+bias) (the tensor the dgrad / wgrad kernels consume) between the
CUDA path (tf32x3) and the fp64 oracle on the same branch decisions, generator only, L = sum(gen * R)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "terra-gan_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import terra_oracle as O
from tg_b200 import precision as PR
from mvp_gan.src.models.generator import PConvUNet
import gates as GT

H, kind = int(sys.argv[1]), sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "tf32x3"
B, DEV = 2, "cuda"
x, mask = O.make_tiles(30, B, H), O.make_mask(31, B, H, kind)
R = torch.randn(B, 1, H, H, generator=torch.Generator().manual_seed(5))
G = PConvUNet()
G.load_state_dict(O.make_generator_state(1))
G.to(DEV).train()
GT.arm(G)
G._engine.debug = {}
with PR.precision(mode):
    out = G((x * mask).to(DEV), mask.to(DEV))
    (out * R.to(DEV)).sum().backward()
gates = GT.collect(G)
dbl = lambda sd: {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
sd = O._require_grad(dbl(O.make_generator_state(1)))
trace = {}
with O.gate_tape(gates) as tape:
    o = O.pconv_unet((x * mask).double(), mask.double(), sd, True, trace)
    for n, *_ in O.ENC + O.DEC:
        trace[n + ".z"].retain_grad()
    (o * R.double()).sum().backward()
print(GT.summarize(tape))
print("out", ((out.double().cpu() - o).abs().max() / o.abs().max()).item())
for n, _, _, k, s, p in O.ENC + O.DEC:
    z = trace[n + ".z"]
    msum = trace[n + ".msum"]
    r = (k * k / (msum + 1e-8)) * (msum > 0)
    ref = (z.grad * r).permute(0, 2, 3, 1)
    got = G._engine.debug[n + ".gz"][:, 0].double().cpu()
    e = ((got - ref).abs().max() / ref.abs().max()).item()
    zz = (G._trace[n + ".y"].double().cpu().permute(0, 3, 1, 2) - trace[n + ".y"]).abs().max() / trace[n + ".y"].abs().max()
    wk = n + ".input_conv.weight"
    ew = ((getattr(G, n).input_conv.weight.grad.double().cpu() - sd[wk].grad).abs().max() / sd[wk].grad.abs().max()).item()
    eg = ((getattr(G, n).bn.weight.grad.double().cpu() - sd[n + ".bn.weight"].grad).abs().max() / sd[n + ".bn.weight"].grad.abs().max()).item()
    print(f"{n}: y {zz.item():.1e}  gz {e:.1e}  dW {ew:.1e} dgamma {eg:.1e}  valid-frac {float((msum > 0).double().mean()):.3f}")
