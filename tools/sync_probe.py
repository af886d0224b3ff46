"""Which call in a train step synchronises the host with the device? torch.cuda.set_sync_debug_mode('warn') + timing of
every phase of AdversarialStep.run on the host."""
import os, sys, time, warnings, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200")); sys.path.insert(0, ROOT)
from tg_b200.step import AdversarialStep
from tg_b200 import optim as tg_optim
from mvp_gan.src.models.generator import PConvUNet
from mvp_gan.src.models.discriminator import Discriminator
from mvp_gan.src.utils.losses import InpaintingLoss
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
G, D = PConvUNet().to(dev).train(), Discriminator().to(dev).train()
os.environ.setdefault("TERRA_VGG_SEED", "3")
crit = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=dev)
st = AdversarialStep(G, D, crit, tg_optim.Adam(G.parameters(), lr=2e-4, modules=[G]), tg_optim.Adam(D.parameters(), lr=2e-4, modules=[D]))
real = torch.rand(B, 1, 512, 512, device=dev)
mask = (torch.rand(B, 1, 512, 512, device=dev) > 0.2).float()
for _ in range(3):
    st.run(real, mask)
torch.cuda.synchronize()
torch.cuda.set_sync_debug_mode("warn")
with warnings.catch_warnings(record=True) as w:
    warnings.simplefilter("always")
    st.run(real, mask)
torch.cuda.set_sync_debug_mode("default")
for x in w:
    print("SYNC WARNING:", str(x.message)[:200], x.filename, x.lineno)
print(len(w), "sync warnings")
# host time of one step when the GPU is not the bottleneck (tiny kernels would be needed; instead: time to ENQUEUE)
torch.cuda.synchronize()
t0 = time.perf_counter()
st.run(real, mask)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue time of one step {1e3*(t1-t0):.2f} ms; until GPU done {1e3*(t2-t0):.2f} ms")
