"""Integration against the REFERENCE'S OWN CALLERS (only where /root/reference exists, i.e. in the build container):
the reference's inference entry mvp_gan/src/evaluate.py is loaded from its source file — unmodified — with this repo's
drop-in `PConvUNet` substituted for `.models.generator`, and executed on real PNG files.

Without a GPU (the build container) everything the caller does around the kernels runs — image decode, Resize,
ToTensor, mask binarisation, `isinstance(model, PConvUNet)`, `PConvUNet().to(device)`, `torch.load`, both checkpoint
forms of `load_state_dict`, `.eval()`, `torch.no_grad()` — and the first kernel call raises the no-CPU-fallback error,
which is the contract. With a GPU (and the reference present) the call completes and the PNG it writes is compared with
the oracle pipeline."""
import importlib.util
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import terra_oracle as O

REF = os.environ.get("TERRA_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "mvp_gan/src/evaluate.py")),
                                reason="the reference tree is only present in the build container")


def load_reference_evaluate():
    import mvp_gan.src.models.generator as ours
    pkg = types.ModuleType("refsrc")
    pkg.__path__ = []
    models = types.ModuleType("refsrc.models")
    models.__path__ = []
    sys.modules.update({"refsrc": pkg, "refsrc.models": models, "refsrc.models.generator": ours})
    spec = importlib.util.spec_from_file_location("refsrc.evaluate", os.path.join(REF, "mvp_gan/src/evaluate.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["refsrc.evaluate"] = mod
    spec.loader.exec_module(mod)          # `from .models.generator import PConvUNet` resolves to the drop-in
    return mod, ours


def test_reference_evaluate_py_drives_the_drop_in_generator(tmp_path):
    from PIL import Image
    ref_eval, ours = load_reference_evaluate()
    assert ref_eval.PConvUNet is ours.PConvUNet
    img = (O.make_tiles(8, 1, 512)[0, 0].numpy() * 255).astype(np.uint8)
    msk = (O.make_mask(9, 1, 512, "large")[0, 0].numpy() * 255).astype(np.uint8)
    ip, mp, op = tmp_path / "tile.png", tmp_path / "tile_mask_resized.png", tmp_path / "tile_inpainted.png"
    Image.fromarray(img, mode="L").save(ip)
    Image.fromarray(msk, mode="L").save(mp)
    sd = O.make_generator_state(1)
    ck_dict, ck_bare = tmp_path / "master_checkpoint.pth", tmp_path / "bare.pth"
    torch.save({"generator_state_dict": sd, "epoch": 3}, ck_dict)      # train.py:318-330 form
    torch.save(sd, ck_bare)                                              # bare state_dict form (evaluate.py:44-45)
    G = ours.PConvUNet()
    G.load_state_dict(sd)
    if torch.cuda.is_available():
        from oracle import image_io as IO
        G.cuda()
        for model in (G, str(ck_dict), str(ck_bare)):
            ref_eval.evaluate(ip, mp, model, op)
            got = np.asarray(Image.open(op))
            image = torch.from_numpy(img).float().div(255)[None, None]
            mask = (torch.from_numpy(msk).float().div(255) > 0).float()[None, None]
            with torch.no_grad():
                want = O.pconv_unet(image * mask, mask, O.make_generator_state(1), False)
            want = IO.pil_resize_bilinear_u8(IO.quantize_u8(want[0, 0].numpy()), 500, 500)
            assert got.shape == (500, 500) and np.abs(got.astype(int) - want.astype(int)).max() <= 2
    else:
        for model in (G, str(ck_dict), str(ck_bare)):
            with pytest.raises(RuntimeError, match="no CPU fallback"):
                ref_eval.evaluate(ip, mp, model, op)
        assert not G.training                                            # the caller's generator.eval() reached the module
