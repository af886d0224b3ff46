"""CUDA-graph replay of generator inference for small batches.

The reference's inference path (mvp_gan/src/evaluate.py:47-50, main_pipeline.py:513-530) calls
`generator(masked, mask)` under eval() / no_grad once per tile (batch 1); BASELINE.json configs[1] batches it
at 16. At those sizes the ~70 kernel launches of one forward cost more host time than GPU time (round 1:
22.6 ms per B=16 call against ~18 ms of kernels). Every launch of the path goes through the C ABI with explicit
workspaces and no host synchronisation, so a whole forward is capturable: `GraphedGenerator` records it once for
a fixed input shape and replays it with one `cudaGraphLaunch` per call.
"""
from __future__ import annotations

import torch


class GraphedGenerator:
    """`out = gg(masked, mask)` == `G.eval(); with no_grad: G(masked, mask)` for inputs of the captured shape.

    The returned tensor is the graph's static output buffer: consume (or copy) it before the next call. Weights
    are read through the engine's packed copies, so re-capture (`gg.capture()`) after the parameters change."""

    def __init__(self, generator, batch: int, height: int, width: int, device=None):
        self.G = generator
        dev = torch.device(device) if device is not None else next(generator.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("GraphedGenerator: the TERRA-GAN B200 path runs on CUDA only (no CPU fallback)")
        self.x = torch.zeros((batch, 1, height, width), device=dev)
        self.mask = torch.ones((batch, 1, height, width), device=dev)
        self.graph = None
        self.out = None
        self.capture()

    def capture(self) -> None:
        if self.G.training:
            raise RuntimeError("GraphedGenerator: put the generator in eval() mode first (evaluate.py:47)")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):                      # builds packed weights, LUT tensors, kernel attributes
                self.G(self.x, self.mask)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.out = self.G(self.x, self.mask)

    def __call__(self, masked: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        if masked.shape != self.x.shape or mask.shape != self.mask.shape:
            raise RuntimeError(f"GraphedGenerator: captured for {tuple(self.x.shape)}, got {tuple(masked.shape)}")
        self.x.copy_(masked, non_blocking=True)
        self.mask.copy_(mask, non_blocking=True)
        self.graph.replay()
        return self.out
