"""Pins the oracle (oracle/terra_oracle.py) to the fixtures that tests/golden/make_golden.py produced
by running the reference's own modules. fp32 on both sides, same seeded inputs: tolerances only
absorb summation-order noise (the restatement calls the same ATen ops)."""
import os

import numpy as np
import pytest
import torch

from oracle import terra_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PCONV_CASES = [("enc1_cfg1", 1, 64, 7, 2, 3, 1, 64, "iid"), ("enc2", 64, 128, 5, 2, 2, 2, 16, "rect"),
               ("enc4", 256, 512, 3, 2, 1, 2, 8, "large"), ("dec2", 192, 64, 3, 1, 1, 2, 16, "rect"),
               ("dec1", 64, 64, 3, 1, 1, 1, 32, "large")]


def close(a, b, rtol=2e-4, atol=2e-5):
    np.testing.assert_allclose(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64), rtol=rtol, atol=atol)


def check_fp(z, key, t, rtol=5e-4):
    t = t.detach().double().reshape(-1)
    ns = z[key + "/norm_sum"]
    assert abs(t.norm().item() - ns[0]) <= rtol * max(ns[0], 1e-6), key
    vals = t[torch.from_numpy(z[key + "/idx"])].numpy()
    scale = max(np.abs(z[key + "/vals"]).max(), ns[0] / max(t.numel(), 1) ** 0.5, 1e-12)
    assert np.abs(vals - z[key + "/vals"]).max() <= 5e-3 * scale, key


@pytest.mark.parametrize("i", range(len(PCONV_CASES)))
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_pconv_layer(i, mode):
    z = np.load(os.path.join(G, "pconv_layers.npz"))
    tag, cin, cout, k, s, p, B, H, kind = PCONV_CASES[i]
    sd = O.make_pconv_state(100 + i, cin, cout, k)
    x = torch.randn((B, cin, H, H), generator=torch.Generator().manual_seed(200 + i)).requires_grad_(True)
    mask = O.make_mask(300 + i, B, H, kind)
    names = ["input_conv.weight", "input_conv.bias", "bn.weight", "bn.bias"]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    y, m = O.pconv2d(x, mask, sd, "", s, p, mode == "train")
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(400 + i))
    grads = torch.autograd.grad(y, [x] + [sd[n] for n in names], gy)
    pre = f"{tag}/{mode}/"
    # masks: bit-exact
    assert np.array_equal(np.packbits(m.numpy().astype(np.uint8)), z[pre + "mask"])
    close(y.detach(), z[pre + "y"])
    close(grads[0], z[pre + "dx"], atol=1e-4)
    if z[pre + "dw"].size > 1:
        close(grads[1], z[pre + "dw"], atol=2e-4)
    check_fp(z, pre + "dw_fp", grads[1])
    close(grads[2], z[pre + "db"], atol=2e-4)
    close(grads[3], z[pre + "dgamma"], atol=2e-4)
    close(grads[4], z[pre + "dbeta"], atol=2e-4)
    close(sd["bn.running_mean"], z[pre + "running_mean"])
    close(sd["bn.running_var"], z[pre + "running_var"])


@pytest.mark.parametrize("mi,kind", list(enumerate(("rect", "large", "iid"))))
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_generator_forward(mi, kind, mode):
    z = np.load(os.path.join(G, "generator.npz"))
    H, B = 128, 2
    x = O.make_tiles(10 + mi, B, H)
    mask = O.make_mask(20 + mi, B, H, kind)
    sd = O.make_generator_state(1)
    trace = {}
    with torch.no_grad():
        y = O.pconv_unet(x * mask, mask, sd, mode == "train", trace)
    pre = f"{kind}/{mode}/"
    close(y, z[pre + "out"], atol=5e-5)
    for name, *_ in O.ENC + O.DEC:
        if mode == "train":
            assert np.array_equal(np.packbits(trace[name + ".mask"].numpy().astype(np.uint8)), z[pre + f"mask_out/{name}"]), name
            if name != "enc1":
                min_ = trace[name + ".in_mask"] if name.startswith("dec") else trace[f"enc{int(name[3]) - 1}.mask"]
                assert np.array_equal(np.packbits(min_.numpy().astype(np.uint8)), z[pre + f"mask_in/{name}"]), name
        check_fp(z, pre + f"feat/{name}", trace[name + ".y"])
    if mode == "train":
        for name in ("enc1", "enc7", "dec7", "dec1"):
            close(sd[name + ".bn.running_mean"], z[pre + f"bn_mean/{name}"], atol=1e-5)
            close(sd[name + ".bn.running_var"], z[pre + f"bn_var/{name}"], atol=1e-5)


def test_adversarial_step():
    z = np.load(os.path.join(G, "adv_step.npz"))
    H, B = 128, 2
    real = O.make_tiles(30, B, H)
    masks = O.make_mask(31, B, H, "rect")
    g_sd, d_sd, vgg = O.make_generator_state(1), O.make_discriminator_state(2), O.make_vgg_state(3)
    opt = {}
    r = O.adversarial_step(real, masks, g_sd, d_sd, vgg, lr=2e-4, opt_state=opt)
    close(r["gen"], z["gen"], atol=5e-5)
    for k in ("g_loss", "g_adv", "g_total", "d_loss", "real_loss", "fake_loss"):
        close(r[k], z["loss/" + k], rtol=1e-4)
    for k in ("l1", "tv", "boundary"):
        close(r["terms"][k], z["loss/" + k], rtol=1e-4)
    for k, g in r["g_grads"].items():
        check_fp(z, "g_grad/" + k, g)
    for k, g in r["d_grads"].items():
        check_fp(z, "d_grad/" + k, g)
    for k, v in g_sd.items():
        if "running_" in k:
            close(v, z["g_buf/" + k], atol=1e-5)
        elif v.is_floating_point() and "mask_conv" not in k:
            check_fp(z, "g_param_after/" + k, v, rtol=1e-5)
    for k, v in d_sd.items():
        if "running_" in k:
            close(v, z["d_buf/" + k], atol=1e-5)   # D's BN statistics advance three times per step
        elif v.is_floating_point():
            check_fp(z, "d_param_after/" + k, v, rtol=1e-5)


def test_human_guided_step():
    z = np.load(os.path.join(G, "hg_step.npz"))
    H, B = 128, 2
    images = O.make_tiles(40, B, H)
    masks = O.make_mask(41, B, H, "large")
    human = 1 - O.make_mask(42, B, H, "rect")
    g_sd, vgg = O.make_generator_state(1), O.make_vgg_state(3)
    r = O.human_guided_step(images, masks, human, g_sd, vgg, lr=1e-4, opt_state={})
    close(r["gen"], z["gen"], atol=5e-5)
    close(r["loss"], z["loss"], rtol=1e-4)
    for k, g in r["g_grads"].items():
        check_fp(z, "g_grad/" + k, g)
    for k, v in g_sd.items():
        if v.is_floating_point() and "mask_conv" not in k and "running_" not in k:
            check_fp(z, "g_param_after/" + k, v, rtol=1e-5)
