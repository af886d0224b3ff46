"""Diagnostic: which side of a tf32x3-vs-oracle gradient difference is the error on? Runs the oracle in fp64 with the
CUDA path's branch decisions and compares (CUDA tf32x3, oracle fp32) against it. Usage: python tools/diag_precision.py H kind"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "terra-gan_b200"), os.path.join(ROOT, "tests")]
import torch
from oracle import terra_oracle as O
from tg_b200 import precision as PR
from mvp_gan.src.models.generator import PConvUNet
from mvp_gan.src.models.discriminator import Discriminator
from mvp_gan.src.utils.losses import InpaintingLoss
import gates as GT
import test_precision_gpu as T

H, kind = int(sys.argv[1]), sys.argv[2]
B = 2
DEV = "cuda"
real, masks = O.make_tiles(30, B, H), O.make_mask(31, B, H, kind)
vgg = O.make_vgg_state(3)
G, D, _ = T._modules()
crit = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=torch.device(DEV), vgg_state_dict=vgg)
GT.arm(G, D, crit)
with PR.precision(sys.argv[3] if len(sys.argv) > 3 else "tf32x3"):
    got = T._run_adversarial(G, D, crit, real.to(DEV), masks.to(DEV))
gates = GT.collect(G, D, crit)
gates["sign.pixel"] = torch.sign(got["gen"].cpu() - real)
dbl = lambda sd: {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
with O.gate_tape(gates):
    r32 = O.adversarial_step(real, masks, O.make_generator_state(1), O.make_discriminator_state(2), vgg)
with O.gate_tape(gates):
    r64 = O.adversarial_step(real.double(), masks.double(), dbl(O.make_generator_state(1)), dbl(O.make_discriminator_state(2)), dbl(vgg))
rows = []
for k in r64["g_grads"]:
    ref = r64["g_grads"][k]
    sc = ref.abs().max().item()
    comp = T._bn_companion(k)
    if comp in r64["g_grads"]:
        sc = max(sc, r64["g_grads"][comp].abs().max().item())
    e_cuda = (got["g_grads"][k].double().cpu() - ref).abs().max().item() / sc
    e_o32 = (r32["g_grads"][k].double() - ref).abs().max().item() / sc
    rows.append((k, e_cuda, e_o32))
rows.sort(key=lambda r: -max(r[1], r[2]))
print(f"{H} {kind}: gradient error vs the fp64 oracle (same branch decisions): name, CUDA, oracle-fp32")
for r in rows[:12]:
    print(f"  {r[0]:28s} {r[1]:.2e} {r[2]:.2e}")
print("gen: cuda", ((got["gen"].double().cpu() - r64["gen"]).abs().max() / r64["gen"].abs().max()).item(),
      "oracle32", ((r32["gen"].double() - r64["gen"]).abs().max() / r64["gen"].abs().max()).item())
