"""Measure the SURVEY §8f kernels on the device (CUDA events) against their HBM roofline, with the reference's CPU path
timed beside them: fused logging metrics, batched inference I/O (uint8 in -> uint8 500x500 out), DSM normalisation +
resize, synthetic mask generation. One JSON line -> stdout."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "terra-gan_b200")]
import numpy as np
import torch
from oracle import terra_oracle as O, image_io as IO, maskgen as OMG
from tg_b200 import ops, maskgen
from tg_b200.inference import BatchedInpainter
from mvp_gan.src.models.generator import PConvUNet

dev = "cuda"
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

def gpu_us(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return min(ts)

out = {"hbm_peak_gbs": peak}
# ---- logging-interval metrics, batch 64 of 512^2 (train.py:229-266)
B, H = 64, 512
pred, target = torch.rand(B, 1, H, H, device=dev), torch.rand(B, 1, H, H, device=dev)
mask = (torch.rand(B, 1, H, H, device=dev) > 0.3).float()
us = gpu_us(lambda: ops.quality_metrics(pred, target, mask))
nbytes = 12 * B * H * H
pc, tc, mc = pred[:4].cpu(), target[:4].cpu(), mask[:4].cpu()
t0 = time.perf_counter(); O.quality_metrics(pc, tc, mc); cpu_s = (time.perf_counter() - t0) * (B / 4)
out["quality_metrics_b64"] = {"us": us, "gbs": nbytes / us / 1e3, "frac_of_hbm_peak": nbytes / us / 1e3 / peak,
                              "reference_cpu_ms_extrapolated_from_4_tiles": cpu_s * 1e3}
# ---- resize 512 -> 500 of 64 tiles (fp32 source quantised on the fly), u8 prepare
x = torch.rand(B, H, H, device=dev)
us = gpu_us(lambda: ops.resize_bilinear_u8(x, (500, 500)))
nb = B * (H * H * 4 + H * 500 * 2 + 500 * 500)
out["quantize_resize_512_to_500_b64"] = {"us": us, "gbs": nb / us / 1e3, "frac_of_hbm_peak": nb / us / 1e3 / peak}
iu = torch.randint(0, 256, (B, H, H), dtype=torch.uint8, device=dev)
us = gpu_us(lambda: ops.u8_prepare(iu, iu))
nb = B * H * H * 10
out["u8_prepare_b64"] = {"us": us, "gbs": nb / us / 1e3, "frac_of_hbm_peak": nb / us / 1e3 / peak}
# ---- batched inference end to end: uint8 host -> uint8 host, 64 tiles, batch 16
G = PConvUNet().to(dev).eval()
inp = BatchedInpainter(G, batch=16)
imgs = torch.randint(0, 256, (64, H, H), dtype=torch.uint8).pin_memory()
msks = (torch.rand(64, H, H) > 0.2).to(torch.uint8).mul(255).pin_memory()
inp(imgs, msks); torch.cuda.synchronize()
t0 = time.perf_counter(); inp(imgs, msks); dt = time.perf_counter() - t0
out["batched_inpainter_u8_to_u8"] = {"tiles_per_s": 64 / dt, "batch": 16, "tiles": 64}
t0 = time.perf_counter()
sd = O.make_generator_state(1)
with torch.no_grad():
    im = imgs[:2].float().div(255).unsqueeze(1); mk = (msks[:2] > 0).float().unsqueeze(1)
    o = O.pconv_unet(im * mk, mk, sd, False)
[IO.pil_resize_bilinear_u8(IO.quantize_u8(t[0].numpy()), 500, 500) for t in o]
out["batched_inpainter_u8_to_u8"]["reference_cpu_tiles_per_s"] = 2 / (time.perf_counter() - t0)
# ---- DSM normalisation + resize to 512 (data_extraction.py:80-107), 16 tiles of 1000x1000
d = torch.rand(16, 1000, 1000, dtype=torch.float64, device=dev) * 300
us = gpu_us(lambda: ops.resize_bilinear_u8(ops.dsm_normalize(d)[0], (512, 512)))
dn = d[0].cpu().numpy()
t0 = time.perf_counter(); IO.pil_resize_bilinear_u8(IO.normalize_dsm(dn), 512, 512); cpu_ms = (time.perf_counter() - t0) * 1e3 * 16
out["dsm_normalize_resize_16x1000x1000"] = {"us": us, "gbs": 16 * 1000 * 1000 * 17 / us / 1e3, "reference_numpy_ms": cpu_ms}
# ---- synthetic masks, size 512
for ap in ("edge", "patch", "region"):
    np.random.seed(11); maskgen.generate_dem_random_mask(512, ap); torch.cuda.synchronize()
    np.random.seed(11); t0 = time.perf_counter(); [maskgen.generate_dem_random_mask(512, ap) for _ in range(4)]; torch.cuda.synchronize(); g = (time.perf_counter() - t0) / 4
    np.random.seed(11); t0 = time.perf_counter(); [OMG.generate_dem_random_mask(512, ap) for _ in range(4)]; c = (time.perf_counter() - t0) / 4
    out[f"mask_{ap}_512"] = {"device_ms_per_mask": g * 1e3, "scipy_cpu_ms_per_mask": c * 1e3}
print(json.dumps(out))
