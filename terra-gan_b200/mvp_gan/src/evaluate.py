"""Drop-in for the reference's inference entry mvp_gan/src/evaluate.py:8-60 — same function, same arguments, same
files written — plus a batched form.

`evaluate(image_path, mask_path, model_or_checkpoint_path, save_path)` keeps the reference's behaviour bit for bit
on the byte side (PIL 'L' decode, Resize((512,512)), ToTensor, mask binarisation, `*255 -> uint8`, 500x500 bilinear
resize, PNG) but everything between the decoded bytes and the encoded bytes runs on the device
(tg_b200.inference.BatchedInpainter). `evaluate_batch` does the same for lists of paths with one generator forward
per `batch` tiles — the loop main_pipeline.py:513-530 runs one tile at a time.
"""
from typing import Sequence

import numpy as np
import torch
from PIL import Image

from .models.generator import PConvUNet
from tg_b200.inference import BatchedInpainter

_inpainters = {}


def _load_generator(model_or_checkpoint_path, device):
    if isinstance(model_or_checkpoint_path, PConvUNet):
        return model_or_checkpoint_path
    generator = PConvUNet().to(device)
    checkpoint = torch.load(model_or_checkpoint_path, map_location=device)
    if isinstance(checkpoint, dict) and 'generator_state_dict' in checkpoint:
        generator.load_state_dict(checkpoint['generator_state_dict'])
    else:
        generator.load_state_dict(checkpoint)
    return generator


def _inpainter(generator, batch):
    key = (id(generator), batch, tuple(p._version for p in generator.parameters()))
    hit = _inpainters.get(id(generator))
    if hit is None or hit[0] != key:
        _inpainters[id(generator)] = (key, BatchedInpainter(generator, batch=batch, use_graph=batch > 1))
    return _inpainters[id(generator)][1]


def evaluate_batch(image_paths: Sequence, mask_paths: Sequence, model_or_checkpoint_path, save_paths: Sequence,
                   batch: int = 16) -> None:
    """Inpaint many tiles: `evaluate` for lists of paths, `batch` tiles per generator forward."""
    if not torch.cuda.is_available():
        raise RuntimeError("evaluate: the B200 TERRA-GAN path runs hand-written sm_100a CUDA kernels only and has no "
                           "CPU fallback")
    device = torch.device('cuda')
    generator = _load_generator(model_or_checkpoint_path, device)
    generator.eval()
    imgs, msks = [], []
    for ip, mp in zip(image_paths, mask_paths):
        # transforms.Resize on a PIL image IS Image.resize(BILINEAR): done on the host only when sizes differ per file
        im, mk = Image.open(ip).convert('L'), Image.open(mp).convert('L')
        if im.size != (512, 512):
            im = im.resize((512, 512), Image.BILINEAR)
        if mk.size != (512, 512):
            mk = mk.resize((512, 512), Image.BILINEAR)
        imgs.append(np.asarray(im))
        msks.append(np.asarray(mk))
    images = torch.from_numpy(np.stack(imgs)).pin_memory()
    masks = torch.from_numpy(np.stack(msks)).pin_memory()
    out = _inpainter(generator, min(batch, max(len(imgs), 1)))(images, masks).numpy()
    for arr, sp in zip(out, save_paths):
        Image.fromarray(arr, mode='L').save(sp)


def evaluate(image_path, mask_path, model_or_checkpoint_path, save_path):
    """Evaluate a model on a single image (reference signature, evaluate.py:8)."""
    evaluate_batch([image_path], [mask_path], model_or_checkpoint_path, [save_path], batch=1)
    print(f"Inpainted image saved to {save_path}")
