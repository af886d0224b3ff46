"""One adversarial train step bracketed by cudaProfilerStart/Stop for ncu (--profile-from-start off).
usage: python tools/profile_step.py <batch> [warmup]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200")); sys.path.insert(0, ROOT)
from tg_b200.step import AdversarialStep
from mvp_gan.src.models.generator import PConvUNet
from mvp_gan.src.models.discriminator import Discriminator
from mvp_gan.src.utils.losses import InpaintingLoss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
W = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
torch.manual_seed(1)
G, D = PConvUNet(), Discriminator()
G.to(dev).train(); D.to(dev).train()
os.environ.setdefault("TERRA_VGG_SEED", "3")
crit = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=dev)
from tg_b200 import optim as tg_optim
st = AdversarialStep(G, D, crit, tg_optim.Adam(G.parameters(), lr=2e-4, modules=[G]),
                     tg_optim.Adam(D.parameters(), lr=2e-4, modules=[D]))     # the bench default: fused Adam + re-pack
gen = torch.Generator().manual_seed(0)
real = torch.rand((B, 1, 512, 512), generator=gen).to(dev)
mask = torch.ones(B, 1, 512, 512)
for b in range(B):
    for _ in range(3):
        hh, ww = (int(torch.randint(32, 257, (1,), generator=gen)) for _ in range(2))
        y0 = int(torch.randint(0, 512 - hh + 1, (1,), generator=gen)); x0 = int(torch.randint(0, 512 - ww + 1, (1,), generator=gen))
        mask[b, 0, y0:y0 + hh, x0:x0 + ww] = 0
mask = mask.to(dev)
for _ in range(W):
    st.run(real, mask)
torch.cuda.synchronize()
torch.cuda.profiler.start()
st.run(real, mask)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
