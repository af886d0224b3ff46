"""How much of a B=64 adversarial step is the GPU idle between kernels? torch.profiler (CUPTI) kernel timeline of 3 steps."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200")); sys.path.insert(0, ROOT)
from tg_b200.step import AdversarialStep
from tg_b200 import optim as tg_optim
from mvp_gan.src.models.generator import PConvUNet
from mvp_gan.src.models.discriminator import Discriminator
from mvp_gan.src.utils.losses import InpaintingLoss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(1)
G, D = PConvUNet().to(dev).train(), Discriminator().to(dev).train()
os.environ.setdefault("TERRA_VGG_SEED", "3")
crit = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=dev)
st = AdversarialStep(G, D, crit, tg_optim.Adam(G.parameters(), lr=2e-4, modules=[G]), tg_optim.Adam(D.parameters(), lr=2e-4, modules=[D]))
real = torch.rand(B, 1, 512, 512, device=dev)
mask = (torch.rand(B, 1, 512, 512, device=dev) > 0.2).float()
for _ in range(3):
    st.run(real, mask)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        st.run(real, mask)
    torch.cuda.synchronize()
evs = sorted([(e.time_range.start, e.time_range.end, e.name) for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start], key=lambda x: x[0])
busy = 0.0
cur_s, cur_e = evs[0][0], evs[0][1]
gaps = []
for s, e, n in evs[1:]:
    if s > cur_e:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, n))
        cur_s, cur_e = s, e
    else:
        cur_e = max(cur_e, e)
busy += cur_e - cur_s
span = evs[-1][1] - evs[0][0]
print(f"B={B}: span {span/3e3:.2f} ms/step, busy {busy/3e3:.2f} ms/step, idle {(span-busy)/3e3:.2f} ms/step over {len(evs)//3} kernels/step")
# context of the largest gap
big = max(range(1, len(evs)), key=lambda i: evs[i][0] - max(e for _, e, _ in evs[max(0, i - 4):i]))
print("around the largest gap:")
for s_, e_, n_ in evs[max(0, big - 4):big + 3]:
    print(f"   start {s_ - evs[0][0]:10.1f} us  dur {e_ - s_:8.1f} us  {n_[:70]}")
gaps.sort(key=lambda g: -g[0])
print("largest gaps (us, next kernel):", [(round(g, 1), n[:40]) for g, n in gaps[:12]])
