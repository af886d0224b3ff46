"""CPU checks of the round-2 "next" rows: the oracle restatements of the logging-interval
metrics and of the byte-level image I/O are pinned to fixtures produced by the reference's own functions and by
Pillow (tests/golden/make_golden.py aux), and the library's host-side coefficient builder equals Pillow's tables."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import image_io as IO
from oracle import terra_oracle as O
from tg_b200 import _lib

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIELDS = ("psnr", "ssim", "l1_distance", "l2_distance", "mse", "boundary_mse", "boundary_psnr", "boundary_gradient_diff")


def metric_case(i, z):
    kind = str(z["kinds"][i])
    B, H, W = (int(v) for v in z[f"{i}/case"])
    target = O.make_tiles(500 + i, B, H, W)
    mask = O.make_mask(510 + i, B, H, kind, W)
    noise = torch.rand((B, 1, H, W), generator=torch.Generator().manual_seed(520 + i))
    pred = target * mask + (0.8 * target + 0.2 * noise) * (1 - mask)
    return pred, target, mask


@pytest.mark.parametrize("i", range(4))
def test_oracle_quality_metrics_match_reference(i):
    z = np.load(os.path.join(G, "metrics.npz"))
    pred, target, mask = metric_case(i, z)
    got = O.quality_metrics(pred, target, mask)
    for name, ref in zip(FIELDS, z[f"{i}/values"]):
        if np.isinf(ref):
            assert np.isinf(got[name]), name
        else:
            assert abs(got[name] - ref) <= 1e-6 * max(abs(ref), 1.0), (name, got[name], ref)


def test_oracle_pillow_resize_quantize_dsm_bit_exact():
    z = np.load(os.path.join(G, "image_io.npz"))
    for i in range(4):
        a, r = z[f"resize/{i}/in"], z[f"resize/{i}/out"]
        assert np.array_equal(IO.pil_resize_bilinear_u8(a, r.shape[0], r.shape[1]), r), i
    assert np.array_equal(IO.quantize_u8(z["quant/in"]), z["quant/out"])
    for i in range(2):
        n = IO.normalize_dsm(z[f"dsm/{i}/in"])
        assert np.array_equal(n, z[f"dsm/{i}/norm"])
        assert np.array_equal(IO.pil_resize_bilinear_u8(n, 512, 512), z[f"dsm/{i}/png"])
    assert not IO.normalize_dsm(np.full((4, 4), np.nan)).any() and not IO.normalize_dsm(np.ones((4, 4))).any()


@pytest.mark.parametrize("in_size,out_size", [(512, 500), (500, 512), (64, 50), (37, 512), (1024, 512)])
def test_library_resize_tables_equal_pillow(in_size, out_size):
    """tg_resize_coeffs runs on the host: the fixed-point tables the kernels consume equal the restated Pillow ones."""
    lib = _lib.lib()
    ks = lib.tg_resize_ksize(in_size, out_size)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ks), dtype=np.int32)
    rc = lib.tg_resize_coeffs(in_size, out_size, bounds.ctypes.data_as(ctypes.c_void_p), kk.ctypes.data_as(ctypes.c_void_p), ks)
    assert rc == 0
    rb, rk = IO._coeffs(in_size, out_size)
    assert ks == rk.shape[1] and np.array_equal(bounds, rb) and np.array_equal(kk, rk)


MASK_CASES = 8


@pytest.mark.parametrize("i", range(MASK_CASES))
def test_oracle_mask_generator_matches_reference(i):
    """oracle.maskgen reproduces the reference's generate_dem_random_mask under the same seed bit for bit."""
    from oracle import maskgen as MG
    z = np.load(os.path.join(G, "masks.npz"))
    seed, size, count = (int(v) for v in z[f"{i}/meta"])
    approach = str(z[f"{i}/approach"])
    np.random.seed(seed)
    m = MG.generate_dem_random_mask(size, None if approach == "none" else approach)
    ref = np.unpackbits(z[f"{i}/bits"])[: size * size].reshape(size, size).astype(bool)
    assert m.sum() == count and np.array_equal(m, ref)
