"""The packed-byte window sum of csrc/mask_pyramid.cu (`mask_window_sum4_kernel`), restated in Python word arithmetic and
checked against the direct definition of the reference's mask convolution (mvp_gan/src/models/pconv.py:33-36: an all-ones
k x k convolution of the 0/1 mask). Documents why the method is exact: window counts (<= 49) never carry into the
neighbouring byte of a 32-bit word, and a horizontal neighbour is a funnel shift across the word boundary. The kernel
itself is tested on the GPU in test_elementwise_gpu.py::test_mask_window_sum_shapes."""
import numpy as np
import pytest

M32 = 0xFFFFFFFF


def funnelshift_l(lo, hi, s):      # upper 32 bits of (hi:lo) << s
    return ((hi << s) | (lo >> (32 - s))) & M32


def funnelshift_r(lo, hi, s):      # lower 32 bits of (hi:lo) >> s
    return ((lo >> s) | (hi << (32 - s))) & M32


def nonzero_bytes(w):              # __vcmpne4(w, 0) & 0x01010101
    return sum(1 << (8 * j) for j in range(4) if (w >> (8 * j)) & 0xFF)


def byte_perm(a, b, sel):
    src = [(a >> (8 * j)) & 0xFF for j in range(4)] + [(b >> (8 * j)) & 0xFF for j in range(4)]
    return sum(src[(sel >> (4 * j)) & 7] << (8 * j) for j in range(4))


def window_sum4(m, k, s):
    """One 'thread' per four consecutive outputs of a row, exactly as the kernel walks them."""
    B, Hi, Wi = m.shape
    pad, Ho, Wo, nw = k // 2, Hi // s, Wi // s, s
    words = m.reshape(B, Hi, Wi // 4, 4).astype(np.uint64)
    wv = words[..., 0] | (words[..., 1] << 8) | (words[..., 2] << 16) | (words[..., 3] << 24)
    out = np.zeros((B, Ho, Wo), np.uint8)
    for b in range(B):
        for ho in range(Ho):
            for q in range(Wo // 4):
                iw0 = q * nw - 1
                V = [0] * (nw + 2)
                for kh in range(k):
                    h = ho * s + kh - pad
                    if 0 <= h < Hi:
                        for j in range(nw + 2):
                            if 0 <= iw0 + j < Wi // 4:
                                V[j] = (V[j] + nonzero_bytes(int(wv[b, h, iw0 + j]))) & M32
                R = []
                for n in range(nw):
                    left, c, right = V[n], V[n + 1], V[n + 2]
                    acc = c
                    for d in range(1, pad + 1):
                        acc = (acc + funnelshift_l(left, c, 8 * d) + funnelshift_r(c, right, 8 * d)) & M32
                    R.append(acc)
                o = R[0] if s == 1 else byte_perm(R[0], R[nw - 1], 0x6420)
                for j in range(4):
                    out[b, ho, 4 * q + j] = (o >> (8 * j)) & 0xFF
    return out


def window_sum_direct(m, k, s):
    B, Hi, Wi = m.shape
    pad, Ho, Wo = k // 2, Hi // s, Wi // s
    mp = np.pad((m != 0).astype(np.int32), ((0, 0), (pad, pad), (pad, pad)))
    out = np.zeros((B, Ho, Wo), np.int32)
    for kh in range(k):
        for kw in range(k):
            out += mp[:, kh:kh + Hi:s, kw:kw + Wi:s][:, :Ho, :Wo]
    return out.astype(np.uint8)


@pytest.mark.parametrize("k,s", [(7, 2), (5, 2), (3, 2), (3, 1)])
@pytest.mark.parametrize("H,W", [(16, 16), (8, 24), (12, 8)])
@pytest.mark.parametrize("fill", ["random", "ones"])
def test_packed_byte_window_sum_is_exact(k, s, H, W, fill):
    rng = np.random.default_rng(k * 10 + s + H)
    if fill == "ones":
        m = np.full((2, H, W), 255, np.uint8)          # largest counts: k*k in the interior, no carry between bytes
    else:
        m = (rng.random((2, H, W)) < 0.6).astype(np.uint8) * rng.integers(1, 255, (2, H, W), dtype=np.uint8)
    got, ref = window_sum4(m, k, s), window_sum_direct(m, k, s)
    assert np.array_equal(got, ref)
    if fill == "ones":
        assert int(ref.max()) == k * k
