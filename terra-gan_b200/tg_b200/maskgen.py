"""Synthetic irregular hole masks on the device.

`generate_dem_random_mask(size, approach)` mirrors the reference's generator of the same name
(/root/reference/random__annotation_mask_generator.py:33-148 — the "edge" / "patch" / "region" artefact masks that
stand in for human annotations): it makes the SAME numpy random draws in the SAME order — so `np.random.seed(s)`
reproduces the reference's mask bit for bit — but every dense image operation (binary morphology, float64 Gaussian
filtering with sigma up to 30, thresholds, distance tests; scipy.ndimage on the CPU in the reference, 0.02-0.2 s per
mask) runs in csrc/maskgen.cu. The mask stays on the device (`torch.bool [size, size]`, True = valid, False = hole)
for the training loop; `hole_masks(n, size, seed)` returns the fp32 [n,1,size,size] batch the loops consume.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import ops
from ._lib import check, lib, ptr, stream_ptr


def _bresenham(x0, y0, x1, y1):
    """Integer line raster (random__annotation_mask_generator.py:9-31); sparse, stays on the host."""
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx, sy = (1 if x0 < x1 else -1), (1 if y0 < y1 else -1)
    err = dx - dy
    pts = []
    while True:
        pts.append((x0, y0))
        if x0 == x1 and y0 == y1:
            return pts
        e2 = 2 * err
        if e2 > -dy:
            err -= dy
            x0 += sx
        if e2 < dx:
            err += dx
            y0 += sy


class _Dev:
    """Device-side image ops on one [size, size] canvas."""

    def __init__(self, size: int, device):
        self.n, self.dev = size, torch.device(device)
        if self.dev.type != "cuda":
            raise RuntimeError("maskgen: the TERRA-GAN B200 path runs on CUDA only (no CPU fallback)")
        self._w = {}

    def morph(self, m: torch.Tensor, erode: bool, iterations: int) -> torch.Tensor:
        for _ in range(iterations):
            out = torch.empty_like(m)
            check(lib().tg_morph_cross(ptr(m), self.n, self.n, 1 if erode else 0, ptr(out), stream_ptr()), "tg_morph_cross")
            m = out
        return m

    def gaussian(self, x: torch.Tensor, sigma: float) -> torch.Tensor:
        """scipy.ndimage.gaussian_filter(x, sigma): truncate 4.0, mode 'reflect', axis 0 then axis 1."""
        radius = int(4.0 * float(sigma) + 0.5)
        xs = np.arange(-radius, radius + 1)
        phi = np.exp(-0.5 / (sigma * sigma) * xs ** 2)
        phi = phi / phi.sum()
        w = torch.from_numpy(phi).to(self.dev)
        for axis in (0, 1):
            out = torch.empty_like(x)
            check(lib().tg_gauss1d_f64(ptr(x), self.n, self.n, axis, ptr(w), radius, ptr(out), stream_ptr()), "tg_gauss1d_f64")
            x = out
        return x

    def shape(self, field, mode, cx, cy, p0, p1, out):
        check(lib().tg_mask_shape(ptr(field), self.n, self.n, mode, int(cx), int(cy), float(p0), float(p1), ptr(out),
                                  stream_ptr()), "tg_mask_shape")


def generate_dem_random_mask(size: int = 500, approach=None, device="cuda") -> torch.Tensor:
    rnd = np.random
    d = _Dev(size, device)
    mask = torch.zeros((size, size), dtype=torch.uint8, device=d.dev)
    if approach is None:
        approach = rnd.choice(["edge", "patch", "region"])
    if approach == "edge":
        base = np.zeros((size, size), dtype=np.uint8)
        for _ in range(rnd.randint(3, 10)):
            pts = rnd.randint(0, size, (rnd.randint(3, 8), 2))
            for j in range(len(pts) - 1):
                for r, c in _bresenham(int(pts[j][0]), int(pts[j][1]), int(pts[j + 1][0]), int(pts[j + 1][1])):
                    if 0 <= r < size and 0 <= c < size:
                        base[r, c] = 1
        b = d.morph(torch.from_numpy(base).to(d.dev), False, rnd.randint(2, 5))
        f = d.gaussian(b.double(), rnd.uniform(1, 3))
        d.shape(f, 0, 0, 0, rnd.uniform(0.4, 0.7), 0.0, mask)
    elif approach == "patch":
        for _ in range(rnd.randint(3, 12)):
            cx, cy = rnd.randint(0, size, 2)
            radius = rnd.randint(10, 50)
            noise = torch.from_numpy(rnd.normal(0, 1, (size, size))).to(d.dev)
            noise = d.gaussian(noise, rnd.uniform(3, 8))
            d.shape(noise, 1, cx, cy, radius, rnd.uniform(5, 15), mask)
    elif approach == "region":
        for _ in range(rnd.randint(1, 4)):
            cx, cy = rnd.randint(0, size, 2)
            min_size = rnd.randint(30, 60)
            max_size = rnd.randint(60, 120)
            if rnd.random() > 0.5:
                a = rnd.randint(min_size, max_size)
                b = rnd.randint(min_size, max_size)
                d.shape(None, 2, cx, cy, a, b, mask)
            else:
                noise = torch.from_numpy(rnd.random((size, size))).to(d.dev)      # row-major fill order of :122-124
                noise = d.gaussian(noise, rnd.uniform(10, 30))
                d.shape(noise, 3, cx, cy, max_size, rnd.uniform(0.4, 0.6), mask)
    else:
        raise ValueError(f"approach must be 'edge', 'patch', 'region' or None, got {approach!r}")
    if rnd.random() > 0.3:                       # binary_opening: erosion then dilation
        it = rnd.randint(1, 2)
        mask = d.morph(d.morph(mask, True, it), False, it)
    if rnd.random() > 0.3:                       # binary_closing: dilation then erosion
        it = rnd.randint(1, 2)
        mask = d.morph(d.morph(mask, False, it), True, it)
    density = float(mask.sum().item()) / (size * size)       # the reference branches on it (:140-144): one host sync
    if density < 0.01:
        mask = d.morph(mask, False, rnd.randint(1, 2))
    elif density > 0.3:
        mask = d.morph(mask, True, rnd.randint(1, 3))
    return mask == 0                             # inverted: True = keep (:146)


def hole_masks(n: int, size: int = 512, seed=None, approach=None, device="cuda") -> torch.Tensor:
    """fp32 [n,1,size,size] batch of masks (1 = valid, 0 = hole) as the training loops consume them."""
    if seed is not None:
        np.random.seed(seed)
    return torch.stack([generate_dem_random_mask(size, approach, device) for _ in range(n)]).unsqueeze(1).float()
