"""Out-of-bounds write detection without compute-sanitizer.

compute-sanitizer is closed on this GPU pool (profiles/r02_compute_sanitizer_closed.txt: "Find a bad access with bounds
checks and asserts of your own, small cases, and a comparison with the CPU reference"). Every output and workspace of
the C-ABI kernels is allocated by tg_b200.ops with torch.empty / empty_like, so this test swaps those allocators for
guarded ones: each buffer sits between two 4 KB canary zones filled with a byte pattern, and after a whole train step,
an inference pass and the auxiliary kernels on RAGGED shapes (odd batch, non-square tiles, partial tiles everywhere) every
canary must be intact. Out-of-bounds reads are covered the usual way: results equal the oracle's (the parity tests)."""
import math

import pytest
import torch

from oracle import terra_oracle as O
from tg_b200 import ops

pytestmark = pytest.mark.gpu
DEV = "cuda"
PAD = 4096
PATTERN = 0xA5


class _GuardedTorch:
    """Stands in for the `torch` module inside tg_b200.ops: empty / empty_like hand out canary-fenced buffers."""

    def __init__(self):
        self.live = []

    def __getattr__(self, name):
        return getattr(torch, name)

    def empty(self, *size, dtype=None, device=None, **kw):
        shape = tuple(size[0]) if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else tuple(size)
        dtype = dtype or torch.float32
        if device is None or torch.device(device).type != "cuda":
            return torch.empty(shape, dtype=dtype, device=device, **kw)
        nbytes = math.prod(shape) * torch.empty((), dtype=dtype).element_size()
        raw = torch.empty((nbytes + 2 * PAD,), dtype=torch.uint8, device=device)
        raw.fill_(PATTERN)
        self.live.append((raw, nbytes, shape, dtype))
        return raw[PAD:PAD + nbytes].view(dtype).reshape(shape)

    def empty_like(self, t, **kw):
        return self.empty(tuple(t.shape), dtype=kw.get("dtype", t.dtype), device=t.device)

    def check(self):
        torch.cuda.synchronize()
        bad = []
        for raw, nbytes, shape, dtype in self.live:
            lo, hi = raw[:PAD], raw[PAD + nbytes:]
            if not (bool((lo == PATTERN).all()) and bool((hi == PATTERN).all())):
                bad.append((shape, dtype, int((lo != PATTERN).sum()), int((hi != PATTERN).sum())))
        return bad


@pytest.fixture
def guard(monkeypatch):
    from tg_b200 import layers, maskgen
    g = _GuardedTorch()
    for mod in (ops, layers, maskgen):          # every module of the package that allocates kernel outputs
        monkeypatch.setattr(mod, "torch", g)
    yield g


@pytest.mark.parametrize("mode", ["bf16", "tf32x3"])
def test_train_step_and_inference_write_inside_their_buffers(guard, mode):
    from tg_b200 import precision as PR
    from tg_b200.step import AdversarialStep
    from mvp_gan.src.models.generator import PConvUNet
    from mvp_gan.src.models.discriminator import Discriminator
    from mvp_gan.src.utils.losses import InpaintingLoss
    B, H, W = 3, 128, 256                                    # odd batch, non-square: partial tiles in every kernel family
    real, masks = O.make_tiles(90, B, H, W).to(DEV), O.make_mask(91, B, H, "large", W).to(DEV)
    G, D = PConvUNet(), Discriminator()
    G.load_state_dict(O.make_generator_state(1))
    D.load_state_dict(O.make_discriminator_state(2))
    G.to(DEV).train()
    D.to(DEV).train()
    crit = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=torch.device(DEV), vgg_state_dict=O.make_vgg_state(3))
    st = AdversarialStep(G, D, crit, torch.optim.Adam(G.parameters(), lr=2e-4), torch.optim.Adam(D.parameters(), lr=2e-4),
                         skip_discarded_d_wgrad=False)
    with PR.precision(mode):
        res = st.run(real, masks)
        G.eval()
        with torch.no_grad():
            out = G(real * masks, masks)
    assert torch.isfinite(out).all() and all(torch.isfinite(v).all() for v in res.values())
    assert len(guard.live) > 300, len(guard.live)           # the guarded allocator really was in the path
    assert guard.check() == []


def test_auxiliary_kernels_write_inside_their_buffers(guard):
    import numpy as np
    from tg_b200 import maskgen
    pred, target = torch.rand(3, 1, 70, 45, device=DEV), torch.rand(3, 1, 70, 45, device=DEV)
    mask = (torch.rand(3, 1, 70, 45, device=DEV) > 0.3).float()
    ops.quality_metrics(pred, target, mask)
    img = torch.randint(0, 256, (3, 37, 41), dtype=torch.uint8, device=DEV)
    ops.resize_bilinear_u8(img, (50, 33))
    ops.resize_bilinear_u8(torch.rand(2, 64, 64, device=DEV), (61, 67))
    ops.u8_prepare(img, img)
    ops.quantize_u8(torch.rand(5, 7, device=DEV))
    ops.dsm_normalize(torch.rand(2, 33, 29, dtype=torch.float64, device=DEV))
    ops.bce_logits_fwd(torch.randn(3, 1, 7, 7, device=DEV), None, 1.0)
    np.random.seed(3)
    maskgen.generate_dem_random_mask(97, "patch")
    assert len(guard.live) >= 10 and guard.check() == []
