"""Generate tests/golden/*.npz by running the REFERENCE's own modules (imported by file path from
/root/reference — the build container only; the GPU box never sees this script run).

The reference has no golden vectors of its own, so these fixtures pin the oracle
(oracle/terra_oracle.py) to the behaviour of the unmodified reference code on seeded inputs:

  * pconv_layers.npz  — single PConv2d layers, fwd + bwd (BASELINE.json config 1 and the other
                        window shapes on the path), train and eval mode
  * generator.npz     — PConvUNet forward at 128x128, B=2, train + eval, every mask of the pyramid
  * adv_step.npz      — one adversarial train step (train.py:179-225) at 128x128, B=2:
                        losses, gradient fingerprints, BN running stats
  * hg_step.npz       — one human-guided step (human_guided_trainer.py:101-153)

Large tensors (gradients of 25.8 M parameters) are stored as fingerprints: L2 norm, sum and 32
values at fixed pseudo-random positions per tensor.

Usage:  python tests/golden/make_golden.py        (needs /root/reference; CPU, ~1 min)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("TERRA_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import terra_oracle as O  # noqa: E402  (only for the shared seeded inputs/weights)


def load_reference():
    """Import pconv/generator/discriminator/losses by path (the mvp_gan package itself imports
    mlflow, which is not installed: SURVEY.md §8c)."""
    pkg = types.ModuleType("refmodels")
    pkg.__path__ = [os.path.join(REF, "mvp_gan/src/models")]
    sys.modules["refmodels"] = pkg
    mods = {}
    for name in ("pconv", "generator", "discriminator"):
        spec = importlib.util.spec_from_file_location(f"refmodels.{name}", os.path.join(REF, f"mvp_gan/src/models/{name}.py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[f"refmodels.{name}"] = m
        spec.loader.exec_module(m)
        mods[name] = m
    spec = importlib.util.spec_from_file_location("reflosses", os.path.join(REF, "mvp_gan/src/utils/losses.py"))
    losses = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(losses)
    mods["losses"] = losses
    return mods


def fingerprint(t: torch.Tensor, n: int = 32):
    t = t.detach().double().reshape(-1)
    g = torch.Generator().manual_seed(t.numel() % 9973 + 17)
    idx = torch.randint(0, t.numel(), (n,), generator=g)
    return np.array([t.norm().item(), t.sum().item()]), idx.numpy(), t[idx].numpy()


def put_fp(out: dict, key: str, t: torch.Tensor):
    ns, idx, vals = fingerprint(t)
    out[key + "/norm_sum"] = ns
    out[key + "/idx"] = idx
    out[key + "/vals"] = vals


def patch_vgg(losses_mod, vgg_sd):
    """vgg16(IMAGENET1K_V1) needs the network; substitute seeded random weights."""
    import torchvision

    def fake_vgg16(weights=None, **kw):
        m = torchvision.models.vgg16(weights=None)
        m.features[:16].load_state_dict(vgg_sd)
        return m
    losses_mod.vgg16 = fake_vgg16


PCONV_CASES = [  # (tag, Cin, Cout, k, stride, pad, B, H, mask kind)
    ("enc1_cfg1", 1, 64, 7, 2, 3, 1, 64, "iid"),
    ("enc2", 64, 128, 5, 2, 2, 2, 16, "rect"),
    ("enc4", 256, 512, 3, 2, 1, 2, 8, "large"),
    ("dec2", 192, 64, 3, 1, 1, 2, 16, "rect"),
    ("dec1", 64, 64, 3, 1, 1, 1, 32, "large"),
]


def gen_pconv_layers(ref):
    out = {}
    for i, (tag, cin, cout, k, s, p, B, H, kind) in enumerate(PCONV_CASES):
        sd = O.make_pconv_state(100 + i, cin, cout, k)
        gen = torch.Generator().manual_seed(200 + i)
        x = torch.randn((B, cin, H, H), generator=gen)
        mask = O.make_mask(300 + i, B, H, kind)
        for mode in ("train", "eval"):
            layer = ref["pconv"].PConv2d(cin, cout, k, s, p)
            layer.load_state_dict(sd)
            layer.train(mode == "train")
            xin = x.clone().requires_grad_(True)
            y, m = layer(xin, mask)
            gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(400 + i))
            y.backward(gy)
            pre = f"{tag}/{mode}/"
            out[pre + "y"] = y.detach().numpy()
            out[pre + "mask"] = np.packbits(m.numpy().astype(np.uint8))
            out[pre + "dx"] = xin.grad.numpy()
            out[pre + "dw"] = layer.input_conv.weight.grad.numpy() if cin * cout * k * k < 70000 else np.zeros(1)
            put_fp(out, pre + "dw_fp", layer.input_conv.weight.grad)
            out[pre + "db"] = layer.input_conv.bias.grad.numpy()
            out[pre + "dgamma"] = layer.bn.weight.grad.numpy()
            out[pre + "dbeta"] = layer.bn.bias.grad.numpy()
            out[pre + "running_mean"] = layer.bn.running_mean.numpy()
            out[pre + "running_var"] = layer.bn.running_var.numpy()
    np.savez_compressed(os.path.join(HERE, "pconv_layers.npz"), **out)
    print("pconv_layers.npz", len(out), "arrays")


def gen_generator(ref):
    out = {}
    H, B = 128, 2
    for mi, kind in enumerate(("rect", "large", "iid")):
        x = O.make_tiles(10 + mi, B, H)
        mask = O.make_mask(20 + mi, B, H, kind)
        for mode in ("train", "eval"):
            G = ref["generator"].PConvUNet()
            G.load_state_dict(O.make_generator_state(1))
            G.train(mode == "train")
            masks = {}
            hooks = []
            for name, mod in G.named_children():
                if name != "final":
                    hooks.append(mod.register_forward_hook(
                        lambda m, i, o, name=name: masks.__setitem__(name, (i[1].detach(), o[1].detach(), o[0].detach()))))
            with torch.no_grad():
                y = G(x * mask, mask)
            for h in hooks:
                h.remove()
            pre = f"{kind}/{mode}/"
            out[pre + "out"] = y.numpy()
            for name, (min_, mout, feat) in masks.items():
                if mode == "train":
                    out[pre + f"mask_in/{name}"] = np.packbits(min_.numpy().astype(np.uint8))
                    out[pre + f"mask_out/{name}"] = np.packbits(mout.numpy().astype(np.uint8))
                put_fp(out, pre + f"feat/{name}", feat)
            if mode == "train":
                for name in ("enc1", "enc7", "dec7", "dec1"):
                    out[pre + f"bn_mean/{name}"] = getattr(G, name).bn.running_mean.numpy()
                    out[pre + f"bn_var/{name}"] = getattr(G, name).bn.running_var.numpy()
    np.savez_compressed(os.path.join(HERE, "generator.npz"), **out)
    print("generator.npz", len(out), "arrays")


def gen_adv_step(ref):
    out = {}
    H, B = 128, 2
    vgg_sd = O.make_vgg_state(3)
    patch_vgg(ref["losses"], vgg_sd)
    real = O.make_tiles(30, B, H)
    masks = O.make_mask(31, B, H, "rect")
    G = ref["generator"].PConvUNet()
    G.load_state_dict(O.make_generator_state(1))
    D = ref["discriminator"].Discriminator()
    D.load_state_dict(O.make_discriminator_state(2))
    G.train()
    D.train()
    dev = torch.device("cpu")
    criterion = ref["losses"].InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=dev)   # train.py:110-114
    adv = torch.nn.BCEWithLogitsLoss()
    opt_G = torch.optim.Adam(G.parameters(), lr=2e-4)
    opt_D = torch.optim.Adam(D.parameters(), lr=2e-4)
    # ---- verbatim order of train.py:179-219 ----
    masked = real * masks
    opt_G.zero_grad()
    gen = G(masked, masks)
    g_loss = criterion(gen, real, masks)
    fake_validity = D(gen)
    g_adv = adv(fake_validity, torch.ones_like(fake_validity))
    g_total = g_loss + g_adv
    g_total.backward()
    g_grads = {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}
    opt_G.step()
    opt_D.zero_grad()
    real_validity = D(real)
    fake_validity = D(gen.detach())
    real_loss = adv(real_validity, torch.ones_like(real_validity))
    fake_loss = adv(fake_validity, torch.zeros_like(fake_validity))
    d_loss = 0.5 * (real_loss + fake_loss)
    d_loss.backward()
    d_grads = {k: p.grad.clone() for k, p in D.named_parameters()}
    opt_D.step()
    # ---------------------------------------------
    out["gen"] = gen.detach().numpy()
    for k, v in dict(g_loss=g_loss, g_adv=g_adv, g_total=g_total, d_loss=d_loss, real_loss=real_loss,
                     fake_loss=fake_loss).items():
        out["loss/" + k] = np.array(v.item())
    with torch.no_grad():
        out["loss/l1"] = np.array(torch.nn.functional.l1_loss(gen, real).item())
        out["loss/tv"] = np.array(criterion.total_variation_loss(gen * (1 - masks)).item())
        out["loss/boundary"] = np.array(criterion.boundary_loss(gen, real, masks).item())
    for k, g in g_grads.items():
        put_fp(out, "g_grad/" + k, g)
    for k, g in d_grads.items():
        put_fp(out, "d_grad/" + k, g)
    for k, p in G.state_dict().items():
        if "running_" in k:
            out["g_buf/" + k] = p.numpy()
        elif p.is_floating_point() and "mask_conv" not in k:
            put_fp(out, "g_param_after/" + k, p)
    for k, p in D.state_dict().items():
        if "running_" in k:
            out["d_buf/" + k] = p.numpy()
        elif p.is_floating_point():
            put_fp(out, "d_param_after/" + k, p)
    np.savez_compressed(os.path.join(HERE, "adv_step.npz"), **out)
    print("adv_step.npz", len(out), "arrays", {k: float(out[k]) for k in out if k.startswith("loss/")})


def gen_hg_step(ref):
    out = {}
    H, B = 128, 2
    vgg_sd = O.make_vgg_state(3)
    patch_vgg(ref["losses"], vgg_sd)
    images = O.make_tiles(40, B, H)
    masks = O.make_mask(41, B, H, "large")
    human = O.make_mask(42, B, H, "rect")
    human = 1 - human  # human-flagged region = the rectangles
    config = {"training": {"loss_weights": {"perceptual": 0.1, "tv": 0.1, "boundary": 0.5},
                           "modes": {"human_guided": {"human_feedback_weight": 0.3, "base_loss_weight": 0.7,
                                                      "learning_rate": 1e-4}}}}
    G = ref["generator"].PConvUNet()
    G.load_state_dict(O.make_generator_state(1))
    G.train()
    criterion = ref["losses"].HumanGuidedLoss(config, device=torch.device("cpu"))
    opt = torch.optim.Adam(G.parameters(), lr=1e-4)
    gen = G(images * masks, masks)
    loss = criterion(gen, images, masks, {"mask": human})
    opt.zero_grad()
    loss.backward()
    grads = {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None}
    opt.step()
    out["gen"] = gen.detach().numpy()
    out["loss"] = np.array(loss.item())
    for k, g in grads.items():
        put_fp(out, "g_grad/" + k, g)
    for k, p in G.state_dict().items():
        if p.is_floating_point() and "mask_conv" not in k and "running_" not in k:
            put_fp(out, "g_param_after/" + k, p)
    np.savez_compressed(os.path.join(HERE, "hg_step.npz"), **out)
    print("hg_step.npz", len(out), "arrays, loss", float(out["loss"]))


def load_reference_metrics():
    """mvp_gan/src/utils/metrics.py (imports GPUtil / psutil, stubbed) and mvp_gan/src/evaluation/metrics.py by path."""
    for stub in ("GPUtil", "psutil"):
        if stub not in sys.modules:
            try:
                __import__(stub)
            except ImportError:
                sys.modules[stub] = types.ModuleType(stub)
    mods = {}
    for tag, rel in (("utils_metrics", "mvp_gan/src/utils/metrics.py"), ("eval_metrics", "mvp_gan/src/evaluation/metrics.py")):
        spec = importlib.util.spec_from_file_location("ref_" + tag, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        mods[tag] = m
    return mods


def gen_metrics():
    """Logging-interval metrics: PerformanceMetrics.calculate_psnr / _ssim / _l1_l2 and
    calculate_boundary_quality of the reference on seeded tiles."""
    ref = load_reference_metrics()
    PM, cbq = ref["utils_metrics"].PerformanceMetrics, ref["eval_metrics"].calculate_boundary_quality
    out = {}
    cases = [("rect", 2, 128, 128), ("large", 3, 96, 160), ("ones", 1, 64, 64), ("iid", 2, 64, 64)]
    for i, (kind, B, H, W) in enumerate(cases):
        target = O.make_tiles(500 + i, B, H, W)
        mask = O.make_mask(510 + i, B, H, kind, W)
        noise = torch.rand((B, 1, H, W), generator=torch.Generator().manual_seed(520 + i))
        pred = target * mask + (0.8 * target + 0.2 * noise) * (1 - mask)
        l1, l2 = PM.calculate_l1_l2(pred, target)
        bq = cbq(pred, target, mask)
        out[f"{i}/values"] = np.array([PM.calculate_psnr(pred, target), PM.calculate_ssim(pred, target), l1, l2,
                                       torch.nn.functional.mse_loss(pred, target).item(), bq["boundary_mse"],
                                       bq["boundary_psnr"], bq["boundary_gradient_diff"]], dtype=np.float64)
        out[f"{i}/case"] = np.array([B, H, W])
    out["kinds"] = np.array([c[0] for c in cases])
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("metrics.npz", {k: v for k, v in out.items() if k.endswith("values")})


def gen_image_io():
    """Batched inference I/O and DSM normalisation (rank 4): Pillow's own
    Image.resize(..., BILINEAR) on uint8 'L' images (evaluate.py:21-25,53-59, data_extraction.py:105-107) and the
    reference's numpy normalisation (data_extraction.py:80-103) on seeded arrays."""
    from PIL import Image
    rng = np.random.RandomState(7)
    out = {}
    for i, (hin, win, hout, wout) in enumerate([(512, 512, 500, 500), (500, 500, 512, 512), (64, 96, 50, 70), (37, 41, 512, 512)]):
        a = rng.randint(0, 256, (hin, win)).astype(np.uint8)
        if i == 0:
            a[:100] = np.arange(win, dtype=np.uint8)[None, :]       # smooth ramp: rounding ties
        r = np.asarray(Image.fromarray(a, mode="L").resize((wout, hout), Image.BILINEAR))
        out[f"resize/{i}/in"], out[f"resize/{i}/out"] = a, r
    # evaluate.py:53-55: (output * 255).astype('uint8')
    f = rng.rand(64, 64).astype(np.float32)
    f[0, :4] = [0.0, 1.0, 0.999999, 0.5]
    out["quant/in"], out["quant/out"] = f, (f * 255).astype("uint8")
    # data_extraction.py:80-103
    for i, (h, w) in enumerate([(200, 200), (50, 80)]):
        d = rng.uniform(-20, 300, (h, w))
        d[rng.rand(h, w) < 0.05] = np.nan
        dmin, dmax = np.nanmin(d), np.nanmax(d)
        n = np.nan_to_num(255 * (d - dmin) / (dmax - dmin), nan=0).astype(np.uint8)
        png = np.asarray(Image.fromarray(n, mode="L").resize((512, 512), Image.BILINEAR))
        out[f"dsm/{i}/in"], out[f"dsm/{i}/norm"], out[f"dsm/{i}/png"] = d, n, png
    np.savez_compressed(os.path.join(HERE, "image_io.npz"), **out)
    print("image_io.npz", len(out), "arrays")


MASK_CASES = [(42, "edge", 500), (43, "patch", 500), (44, "region", 500), (45, "region", 500), (46, "region", 500),
              (47, None, 500), (48, "edge", 512), (49, "patch", 256)]


def gen_masks():
    """Synthetic irregular hole masks: the reference's generate_dem_random_mask under
    np.random.seed(s) (random__annotation_mask_generator.py:33-148; matplotlib, which that file imports for its
    plotting helpers, is stubbed). Stored bit-packed."""
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    spec = importlib.util.spec_from_file_location("ref_maskgen", os.path.join(REF, "random__annotation_mask_generator.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    out = {}
    for i, (seed, approach, size) in enumerate(MASK_CASES):
        np.random.seed(seed)
        mask = m.generate_dem_random_mask(size=size, approach=approach)
        out[f"{i}/bits"] = np.packbits(mask)
        out[f"{i}/meta"] = np.array([seed, size, int(mask.sum())])
        out[f"{i}/approach"] = np.array(approach or "none")
        print("mask", i, seed, approach, size, "valid fraction", float(mask.mean()))
    np.savez_compressed(os.path.join(HERE, "masks.npz"), **out)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    if len(sys.argv) > 1 and sys.argv[1] == "masks":
        gen_masks()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "aux":      # the round-2 fixtures only
        gen_metrics()
        gen_image_io()
        gen_masks()
        sys.exit(0)
    ref = load_reference()
    gen_pconv_layers(ref)
    gen_generator(ref)
    gen_adv_step(ref)
    gen_hg_step(ref)
    gen_metrics()
    gen_image_io()
