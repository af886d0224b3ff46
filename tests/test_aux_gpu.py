"""GPU parity of the round-2 "next" rows through the C ABI: fused logging-interval metrics,
Pillow-exact batched resize / quantisation, DSM normalisation, and the batched inference pipeline end to end
(drop-in evaluate() writing PNGs) against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import image_io as IO
from oracle import terra_oracle as O
from tg_b200 import ops
from tg_b200.inference import BatchedInpainter
from mvp_gan.src.models.generator import PConvUNet
from mvp_gan.src.utils.metrics import PerformanceMetrics, TrainingMetrics, quality_metrics
from mvp_gan.src.evaluation.metrics import calculate_boundary_quality

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("i", range(4))
def test_quality_metrics_vs_reference_golden(i):
    """One fused launch vs the reference's own PerformanceMetrics / calculate_boundary_quality (golden fixture)."""
    from test_aux_cpu import FIELDS, metric_case
    z = np.load(os.path.join(G, "metrics.npz"))
    pred, target, mask = metric_case(i, z)
    m = quality_metrics(pred.to(DEV), target.to(DEV), mask.to(DEV)).to_dict()
    for name, ref in zip(FIELDS, z[f"{i}/values"]):
        if np.isinf(ref):
            assert np.isinf(m[name]), name
        else:
            assert abs(m[name] - ref) <= 2e-5 * max(abs(ref), 1e-3), (name, m[name], ref)
    # the reference-named entry points read the same launch
    p, t, k = pred.to(DEV), target.to(DEV), mask.to(DEV)
    assert PerformanceMetrics.calculate_psnr(p, t) == m["psnr"] and PerformanceMetrics.calculate_ssim(p, t) == m["ssim"]
    assert PerformanceMetrics.calculate_l1_l2(p, t) == (m["l1_distance"], m["l2_distance"])
    bq = calculate_boundary_quality(p, t, k)
    assert set(bq) == {"boundary_mse", "boundary_psnr", "boundary_gradient_diff"} and bq["boundary_mse"] == m["boundary_mse"]


def test_quality_metrics_full_size_vs_oracle():
    B, H = 4, 512
    target, mask = O.make_tiles(1, B, H), O.make_mask(2, B, H, "large")
    pred = (target + 0.05 * torch.randn(B, 1, H, H, generator=torch.Generator().manual_seed(3))).clamp(0, 1)
    ref = O.quality_metrics(pred, target, mask)
    m = quality_metrics(pred.to(DEV), target.to(DEV), mask.to(DEV)).to_dict()
    for k, v in ref.items():
        assert abs(m[k] - v) <= 2e-5 * max(abs(v), 1e-3), (k, m[k], v)


def test_gradient_norms_one_sync():
    net = torch.nn.Sequential(torch.nn.Conv2d(1, 4, 3), torch.nn.Conv2d(4, 1, 3)).to(DEV)
    net(torch.randn(2, 1, 8, 8, device=DEV)).sum().backward()
    got = TrainingMetrics.calculate_gradient_norm(net)
    tot = 0.0
    for n, p in net.named_parameters():
        assert abs(got[f"grad_norm_{n}"] - p.grad.norm(2).item()) < 1e-5
        tot += p.grad.norm(2).item() ** 2
    assert abs(got["total_grad_norm"] - tot ** 0.5) < 1e-5


def test_resize_quantize_dsm_bit_exact_with_pillow():
    z = np.load(os.path.join(G, "image_io.npz"))
    for i in range(4):
        a, r = z[f"resize/{i}/in"], z[f"resize/{i}/out"]
        batch = torch.from_numpy(np.stack([a, a[::-1].copy()])).to(DEV)
        out = ops.resize_bilinear_u8(batch, r.shape).cpu().numpy()
        assert np.array_equal(out[0], r), i
        assert np.array_equal(out[1], IO.pil_resize_bilinear_u8(a[::-1].copy(), *r.shape)), i
    f = torch.from_numpy(z["quant/in"]).to(DEV)
    assert np.array_equal(ops.quantize_u8(f).cpu().numpy(), z["quant/out"])
    # fp32 source quantised on the fly inside the resize (evaluate.py:54-59)
    x = torch.rand((2, 512, 512), generator=torch.Generator().manual_seed(9))
    want = np.stack([IO.pil_resize_bilinear_u8(IO.quantize_u8(t.numpy()), 500, 500) for t in x])
    assert np.array_equal(ops.resize_bilinear_u8(x.to(DEV), (500, 500)).cpu().numpy(), want)
    for i in range(2):
        d = torch.from_numpy(z[f"dsm/{i}/in"]).unsqueeze(0).to(DEV)
        n, mm = ops.dsm_normalize(d)
        assert np.array_equal(n[0].cpu().numpy(), z[f"dsm/{i}/norm"])
        assert np.array_equal(ops.resize_bilinear_u8(n, (512, 512))[0].cpu().numpy(), z[f"dsm/{i}/png"])
    flat = torch.cat([torch.full((1, 8, 8), float("nan"), dtype=torch.float64), torch.ones((1, 8, 8), dtype=torch.float64)]).to(DEV)
    assert not ops.dsm_normalize(flat)[0].any()


def test_u8_prepare_matches_to_tensor_and_binarise():
    g = torch.Generator().manual_seed(4)
    img = torch.randint(0, 256, (3, 64, 64), dtype=torch.uint8, generator=g)
    msk = torch.randint(0, 3, (3, 64, 64), dtype=torch.uint8, generator=g) * 100
    masked, mask = ops.u8_prepare(img.to(DEV), msk.to(DEV))
    image = img.float().div(255).unsqueeze(1)                  # ToTensor
    m = ((msk.float().div(255)) > 0).float().unsqueeze(1)      # evaluate.py:32
    assert torch.equal(mask.cpu(), m) and torch.equal(masked.cpu(), image * m)


def _oracle_inpaint_u8(images_u8, masks_u8, sd):
    image = torch.from_numpy(images_u8).float().div(255).unsqueeze(1)
    mask = (torch.from_numpy(masks_u8).float().div(255) > 0).float().unsqueeze(1)
    with torch.no_grad():
        out = O.pconv_unet(image * mask, mask, sd, False)
    return np.stack([IO.pil_resize_bilinear_u8(IO.quantize_u8(o[0].numpy()), 500, 500) for o in out])


def test_batched_inpainter_vs_oracle_pipeline():
    """uint8 in -> uint8 500x500 out for 5 tiles with batch 2 (graph chunks + a ragged eager chunk)."""
    N = 5
    rng = np.random.RandomState(5)
    images = (O.make_tiles(6, N, 512)[:, 0].numpy() * 255).astype(np.uint8)
    masks = (O.make_mask(7, N, 512, "rect")[:, 0].numpy() * 255).astype(np.uint8)
    masks[0][masks[0] > 0] = 1                                  # any value > 0 is valid (evaluate.py:32)
    Gm = PConvUNet()
    Gm.load_state_dict(O.make_generator_state(1))
    Gm.to(DEV).eval()
    out = BatchedInpainter(Gm, batch=2)(torch.from_numpy(images), torch.from_numpy(masks)).numpy()
    ref = _oracle_inpaint_u8(images, masks, O.make_generator_state(1))
    assert out.shape == ref.shape == (N, 500, 500) and out.dtype == np.uint8
    diff = np.abs(out.astype(int) - ref.astype(int))
    print("batched inpainter vs oracle: max |diff|", diff.max(), "mean", diff.mean(), "pixels differing", (diff > 0).mean())
    assert diff.max() <= 2 and diff.mean() < 0.05             # bf16 generator: an occasional 1-level step at a rounding edge


def test_evaluate_drop_in_writes_the_reference_png(tmp_path):
    """mvp_gan/src/evaluate.py:8 signature: PNG paths in, 500x500 'L' PNG out; also through utils.gan_inpainting."""
    from PIL import Image
    from mvp_gan.src.evaluate import evaluate
    from utils.gan_inpainting import inpaint_with_gan
    img = (O.make_tiles(8, 1, 512)[0, 0].numpy() * 255).astype(np.uint8)
    msk = (O.make_mask(9, 1, 512, "large")[0, 0].numpy() * 255).astype(np.uint8)
    ip, mp, op = tmp_path / "tile.png", tmp_path / "tile_mask_resized.png", tmp_path / "out.png"
    Image.fromarray(img, mode="L").save(ip)
    Image.fromarray(msk, mode="L").save(mp)
    sd = O.make_generator_state(1)
    ckpt = tmp_path / "master_checkpoint.pth"
    torch.save({"generator_state_dict": sd, "epoch": 1}, ckpt)          # the dict form of train.py:318-330
    evaluate(ip, mp, str(ckpt), op)
    got = np.asarray(Image.open(op))
    ref = _oracle_inpaint_u8(img[None], msk[None], sd)[0]
    assert got.shape == (500, 500) and got.dtype == np.uint8
    d = np.abs(got.astype(int) - ref.astype(int))
    assert d.max() <= 2 and d.mean() < 0.05
    Gm = PConvUNet()
    Gm.load_state_dict(sd)
    Gm.to(DEV)
    out2 = inpaint_with_gan(ip, mp, tmp_path / "inpainted", Gm)           # a model instance is accepted too (:36)
    assert np.array_equal(np.asarray(Image.open(out2)), got) and out2.name == "tile_inpainted.png"


@pytest.mark.parametrize("i", range(8))
def test_device_mask_generator_bit_exact_with_reference(i):
    """tg_b200.maskgen: same seed -> the reference's mask, bit for bit (golden fixture from the reference function),
    with the dense scipy.ndimage work done by csrc/maskgen.cu."""
    from tg_b200 import maskgen
    z = np.load(os.path.join(G, "masks.npz"))
    seed, size, count = (int(v) for v in z[f"{i}/meta"])
    approach = str(z[f"{i}/approach"])
    np.random.seed(seed)
    m = maskgen.generate_dem_random_mask(size, None if approach == "none" else approach).cpu().numpy()
    ref = np.unpackbits(z[f"{i}/bits"])[: size * size].reshape(size, size).astype(bool)
    print(i, approach, size, "pixels differing:", int((m != ref).sum()))
    assert np.array_equal(m, ref)


def test_hole_mask_batch_feeds_the_generator():
    from tg_b200 import maskgen
    masks = maskgen.hole_masks(3, 512, seed=123)
    assert masks.shape == (3, 1, 512, 512) and masks.dtype == torch.float32 and masks.is_cuda
    assert set(masks.unique().tolist()) <= {0.0, 1.0} and 0.3 < masks.mean().item() < 1.0
