"""Batched inference pipeline.

The reference inpaints one PNG at a time (main_pipeline.py:513-530 -> mvp_gan/src/evaluate.py:8-60): PIL decode,
torchvision Resize + ToTensor on the host, H2D, batch-1 generator forward, D2H, numpy `*255 -> uint8`, PIL resize to
500x500, PNG encode. `BatchedInpainter` keeps everything between the decoded bytes and the encoded bytes on the
device and batches it: uint8 tiles and masks travel from pinned host buffers on a copy stream (double-buffered, so
the copy of chunk i+1 overlaps the forward of chunk i), one kernel turns them into the masked fp32 image and mask,
the generator forward is one CUDA-graph launch (tg_b200.graphs), and one kernel pair quantises the output and
applies Pillow's bilinear resize bit-exactly (csrc/image_io.cu), so only `batch x 500 x 500` bytes return to the host.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .graphs import GraphedGenerator


class BatchedInpainter:
    def __init__(self, generator, batch: int = 16, tile: int = 512, out_size: Optional[int] = 500, device=None,
                 use_graph: bool = True):
        self.G = generator
        self.batch, self.tile, self.out_size = batch, tile, out_size
        dev = torch.device(device) if device is not None else next(generator.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("BatchedInpainter: the TERRA-GAN B200 path runs on CUDA only (no CPU fallback)")
        self.dev = dev
        generator.eval()                                       # evaluate.py:47
        self.graph = GraphedGenerator(generator, batch, tile, tile, dev) if use_graph else None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self._dev_in = [None, None]
        self._ready = [torch.cuda.Event(), torch.cuda.Event()]
        self._consumed = [torch.cuda.Event(), torch.cuda.Event()]

    # ---- device part: uint8 [n,H,W] x2 (already on the device) -> uint8 [n,out,out] ----
    def _forward_device(self, img_u8: torch.Tensor, mask_u8: torch.Tensor) -> torch.Tensor:
        n = img_u8.shape[0]
        T = self.tile
        if img_u8.shape[1:] != (T, T):                         # transforms.Resize((512,512)) on the 'L' images (:21-25)
            img_u8 = ops.resize_bilinear_u8(img_u8, (T, T))
            mask_u8 = ops.resize_bilinear_u8(mask_u8, (T, T))
        masked, mask = ops.u8_prepare(img_u8.contiguous(), mask_u8.contiguous())
        if self.graph is not None and n == self.batch:
            out = self.graph(masked, mask)
        else:                                                  # ragged last chunk: eager launches
            with torch.no_grad():
                out = self.G(masked, mask)
        if self.out_size is None or self.out_size == T:
            return ops.quantize_u8(out.reshape(n, T, T).contiguous())
        return ops.resize_bilinear_u8(out.reshape(n, T, T), (self.out_size, self.out_size))

    def __call__(self, images_u8: torch.Tensor, masks_u8: torch.Tensor) -> torch.Tensor:
        """images_u8 / masks_u8: uint8 [N,H,W] host tensors (pinned for full overlap) -> uint8 [N,out,out] pinned host."""
        if images_u8.dtype != torch.uint8 or masks_u8.dtype != torch.uint8 or images_u8.shape != masks_u8.shape:
            raise RuntimeError("BatchedInpainter: images and masks must be uint8 tensors of the same [N,H,W] shape")
        N = images_u8.shape[0]
        osz = self.out_size or self.tile
        result = torch.empty((N, osz, osz), dtype=torch.uint8).pin_memory()
        chunks = [(s, min(s + self.batch, N)) for s in range(0, N, self.batch)]
        cur = torch.cuda.current_stream(self.dev)

        def prefetch(ci):
            s, e = chunks[ci]
            slot = ci & 1
            self.copy_stream.wait_event(self._consumed[slot])
            with torch.cuda.stream(self.copy_stream):
                self._dev_in[slot] = (images_u8[s:e].to(self.dev, non_blocking=True),
                                      masks_u8[s:e].to(self.dev, non_blocking=True))
                self._ready[slot].record()

        if chunks:
            prefetch(0)
        for ci, (s, e) in enumerate(chunks):
            slot = ci & 1
            if ci + 1 < len(chunks):
                prefetch(ci + 1)
            cur.wait_event(self._ready[slot])
            img, msk = self._dev_in[slot]
            img.record_stream(cur)
            msk.record_stream(cur)
            out = self._forward_device(img, msk)
            self._consumed[slot].record(cur)
            result[s:e].copy_(out, non_blocking=True)
        cur.synchronize()
        return result
