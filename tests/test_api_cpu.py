"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol the header
declares, the Python mirror has the reference's module API / state_dict layout, and the hot path
refuses to run without CUDA (no fallback)."""
import ctypes
import os
import pickle
import re

import pytest
import torch

from oracle import terra_oracle as O
from tg_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    hdr = open(os.path.join(ROOT, "include", "terragan_b200.h")).read()
    declared = set(re.findall(r"\b(tg_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/terragan_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes and header declarations differ"
    assert _lib.lib().tg_version() == 2


def test_error_reporting_without_gpu_compute():
    # argument validation happens before any device work: callable on a CPU-only box
    rc = _lib.lib().tg_mask_merge_up(None, None, 1, 3, 3, None, None)
    assert rc != 0 and "null pointer" in _lib.last_error()


def test_state_dict_layout_matches_reference():
    from mvp_gan.src.models.generator import PConvUNet
    from mvp_gan.src.models.discriminator import Discriminator
    G, D = PConvUNet(), Discriminator()
    gref, dref = O.make_generator_state(1), O.make_discriminator_state(2)
    gsd, dsd = G.state_dict(), D.state_dict()
    assert len(gsd) == 114 and len(dsd) == 25                      # SURVEY.md §8b [probe]
    assert list(gsd.keys()) == list(gref.keys())
    assert {k: tuple(v.shape) for k, v in gsd.items()} == {k: tuple(v.shape) for k, v in gref.items()}
    assert set(dsd.keys()) == set(dref.keys())
    assert {k: tuple(v.shape) for k, v in dsd.items()} == {k: tuple(v.shape) for k, v in dref.items()}
    G.load_state_dict(gref)
    D.load_state_dict(dref)
    assert sum(p.numel() for p in G.parameters() if p.requires_grad) == 25_813_953
    assert sum(p.numel() for p in D.parameters()) == 2_764_481
    assert not G.enc1.mask_conv.weight.requires_grad and float(G.enc1.mask_conv.weight.min()) == 1.0
    assert G.enc3.slide_winsize == 25 and G.__class__.__name__ == "PConvUNet"
    assert [n for n, _ in G.named_children()] == [f"enc{i}" for i in range(1, 8)] + [f"dec{i}" for i in range(7, 0, -1)] + ["final"]


def test_modules_are_picklable_and_movable():
    from mvp_gan.src.models.generator import PConvUNet
    G = PConvUNet()
    _ = G._engine                         # create the derived cache; it must not end up in the pickle
    G2 = pickle.loads(pickle.dumps(G))
    assert "_engine_cache" not in G2.__dict__
    assert torch.equal(G2.final.weight, G.final.weight)
    assert isinstance(str(G), str) and "PConv2d" in str(G)
    G.cpu()                               # train.py:341 moves the generator to the CPU for MLflow logging


def test_forward_on_cpu_raises_no_fallback():
    from mvp_gan.src.models.generator import PConvUNet
    from mvp_gan.src.models.discriminator import Discriminator
    from mvp_gan.src.models.pconv import PConv2d
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PConvUNet()(torch.rand(1, 1, 128, 128), torch.ones(1, 1, 128, 128))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Discriminator()(torch.rand(1, 1, 64, 64))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PConv2d(1, 64, 7, 2, 3)(torch.rand(1, 1, 64, 64), torch.ones(1, 1, 64, 64))


def test_loss_module_api():
    from mvp_gan.src.utils.losses import BoundaryAwareLoss, HumanGuidedLoss, InpaintingLoss
    vgg = O.make_vgg_state(3)
    crit = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=torch.device("cpu"), vgg_state_dict=vgg)
    assert crit.boundary_weight == 0.5 and isinstance(crit.boundary_loss, BoundaryAwareLoss)
    assert len(crit.vgg_layers) == 16 and not any(p.requires_grad for p in crit.vgg_layers.parameters())
    assert isinstance(crit.l1_loss, torch.nn.L1Loss)
    cfg = {"training": {"loss_weights": {"boundary": 0.25},
                        "modes": {"human_guided": {"human_feedback_weight": 0.3, "base_loss_weight": 0.7}}}}
    hg = HumanGuidedLoss(cfg, device=torch.device("cpu"), vgg_state_dict=vgg)
    assert hg.boundary_weight == 0.25 and hg.base_loss_weight == 0.7
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        crit(torch.rand(1, 1, 64, 64), torch.rand(1, 1, 64, 64), torch.ones(1, 1, 64, 64))
