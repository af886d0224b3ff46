"""Diagnostic: d(loss term)/d(pred) of every loss branch (fused L1/TV/boundary, VGG perceptual, D + BCE) — CUDA tf32x3
vs the fp64 oracle on the same branch decisions, for a fixed prediction tensor."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "terra-gan_b200"), os.path.join(ROOT, "tests")]
import torch
import torch.nn.functional as F
from oracle import terra_oracle as O
from tg_b200 import precision as PR
from mvp_gan.src.models.discriminator import Discriminator
from mvp_gan.src.utils.losses import InpaintingLoss
import gates as GT

H, kind = int(sys.argv[1]), sys.argv[2]
mode = sys.argv[3] if len(sys.argv) > 3 else "tf32x3"
B, DEV = 2, "cuda"
real, mask = O.make_tiles(30, B, H), O.make_mask(31, B, H, kind)
pred = (torch.rand(B, 1, H, H, generator=torch.Generator().manual_seed(6)) * (1 - mask) + real * mask)
vgg = O.make_vgg_state(3)
dbl = lambda sd: {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
D = Discriminator(); D.load_state_dict(O.make_discriminator_state(2)); D.to(DEV).train()
rel = lambda a, b: ((a.double().cpu() - b).abs().max() / b.abs().max()).item()

def cuda_grad(fn):
    p = pred.to(DEV).requires_grad_(True)
    with PR.precision(mode):
        l = fn(p)
        l.backward()
    return l.item(), p.grad

def ref_grad(fn, gates):
    p = pred.double().requires_grad_(True)
    with O.gate_tape(gates) as tape:
        l = fn(p)
        l.backward()
    return l.item(), p.grad, tape

for name, pw, tw, bw in (("l1", 0, 0, 0), ("tv", 0, 1.0, 0), ("boundary", 0, 0, 1.0), ("perceptual", 1.0, 0, 0)):
    crit = InpaintingLoss(perceptual_weight=pw, tv_weight=tw, boundary_weight=bw, device=torch.device(DEV), vgg_state_dict=vgg)
    GT.arm(None, None, crit)
    lc, gc = cuda_grad(lambda p: crit(p, real.to(DEV), mask.to(DEV)))
    gates = GT.collect(None, None, crit)
    lr, gr, tape = ref_grad(lambda p: O.inpainting_loss(p, real.double(), mask.double(), dbl(vgg), pw, tw, bw), gates)
    print(f"{name:10s} loss {lc:.8f} vs {lr:.8f}  grad err {rel(gc, gr):.2e}  |", GT.summarize(tape))

GT.arm(None, D, None)
lc, gc = cuda_grad(lambda p: F.binary_cross_entropy_with_logits(D(p), torch.ones(B, 1, H // 16 - 1, H // 16 - 1, device=DEV)))
gates = GT.collect(None, D, None)
lr, gr, tape = ref_grad(lambda p: F.binary_cross_entropy_with_logits(O.discriminator(p, dbl(O.make_discriminator_state(2)), True), torch.ones(B, 1, H // 16 - 1, H // 16 - 1, dtype=torch.double)), gates)
print(f"{'adversarial':10s} loss {lc:.8f} vs {lr:.8f}  grad err {rel(gc, gr):.2e}  |", GT.summarize(tape))
