"""TEST INFRASTRUCTURE ONLY: CPU restatement of the reference's synthetic hole-mask generator
`generate_dem_random_mask(size, approach)` — /root/reference/random__annotation_mask_generator.py:33-148 (the spec of
the irregular masks of the human-guided workload). Same numpy RNG draws in the same
order, the same scipy.ndimage calls (scipy is the third-party dependency the algorithm lives in), so with
`np.random.seed(s)` it returns the reference's masks bit for bit — pinned by tests/golden/masks.npz, which
tests/golden/make_golden.py produced by running the reference function itself.
Returned mask: bool [size, size], True = keep (valid), False = hole (the reference inverts at :146)."""
import numpy as np
from scipy import ndimage


def bresenham(x0, y0, x1, y1):
    """Integer line between two points (random__annotation_mask_generator.py:9-31)."""
    dx, dy = abs(x1 - x0), abs(y1 - y0)
    sx, sy = (1 if x0 < x1 else -1), (1 if y0 < y1 else -1)
    err = dx - dy
    rr, cc = [], []
    while True:
        rr.append(x0)
        cc.append(y0)
        if x0 == x1 and y0 == y1:
            break
        e2 = 2 * err
        if e2 > -dy:
            err -= dy
            x0 += sx
        if e2 < dx:
            err += dx
            y0 += sy
    return np.array(rr), np.array(cc)


def generate_dem_random_mask(size=500, approach=None):
    rnd = np.random
    mask = np.zeros((size, size), dtype=bool)
    if approach is None:
        approach = rnd.choice(["edge", "patch", "region"])                       # :47-48
    if approach == "edge":                                                        # :50-76
        base = np.zeros((size, size))
        for _ in range(rnd.randint(3, 10)):
            pts = rnd.randint(0, size, (rnd.randint(3, 8), 2))
            for j in range(len(pts) - 1):
                rr, cc = bresenham(pts[j][0], pts[j][1], pts[j + 1][0], pts[j + 1][1])
                ok = (rr >= 0) & (rr < size) & (cc >= 0) & (cc < size)
                if ok.any():
                    base[rr[ok], cc[ok]] = 1
        base = ndimage.binary_dilation(base, iterations=rnd.randint(2, 5))
        base = ndimage.gaussian_filter(base.astype(float), sigma=rnd.uniform(1, 3))
        mask = base > rnd.uniform(0.4, 0.7)
    elif approach == "patch":                                                     # :78-98
        for _ in range(rnd.randint(3, 12)):
            cx, cy = rnd.randint(0, size, 2)
            radius = rnd.randint(10, 50)
            y, x = np.ogrid[-cy:size - cy, -cx:size - cx]
            dist = np.sqrt(x * x + y * y)
            noise = ndimage.gaussian_filter(rnd.normal(0, 1, (size, size)), sigma=rnd.uniform(3, 8))
            mask = mask | (dist <= radius + noise * rnd.uniform(5, 15))
    elif approach == "region":                                                    # :100-131
        for _ in range(rnd.randint(1, 4)):
            cx, cy = rnd.randint(0, size, 2)
            min_size = rnd.randint(30, 60)
            max_size = rnd.randint(60, 120)
            y, x = np.ogrid[-cy:size - cy, -cx:size - cx]
            if rnd.random() > 0.5:
                a = rnd.randint(min_size, max_size)
                b = rnd.randint(min_size, max_size)
                region = (x * x) / (a * a) + (y * y) / (b * b) <= 1
            else:
                noise = rnd.random((size, size))        # the reference fills it element by element in row-major order (:122-124): same stream
                noise = ndimage.gaussian_filter(noise, sigma=rnd.uniform(10, 30))
                region = ((x * x + y * y) <= max_size ** 2) & (noise > rnd.uniform(0.4, 0.6))
            mask = mask | region
    if rnd.random() > 0.3:                                                        # :134-137
        mask = ndimage.binary_opening(mask, iterations=rnd.randint(1, 2))
    if rnd.random() > 0.3:
        mask = ndimage.binary_closing(mask, iterations=rnd.randint(1, 2))
    density = mask.mean()                                                         # :140-144
    if density < 0.01:
        mask = ndimage.binary_dilation(mask, iterations=rnd.randint(1, 2))
    elif density > 0.3:
        mask = ndimage.binary_erosion(mask, iterations=rnd.randint(1, 3))
    return ~mask                                                                  # :146
