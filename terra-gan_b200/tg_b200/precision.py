"""Arithmetic / storage precision of the hot path.

  "bf16"    product path: bf16 activations, kind::f16 tcgen05 MMAs, fp32 accumulation (the default).
  "tf32"    verification path: fp32 activations, kind::tf32 MMAs (one pass), fp32 everywhere else.
  "tf32x3"  verification path: fp32 activations, every tensor-core product evaluated as hi*hi + lo*hi + hi*lo of
            two-term TF32 splits (three kind::tf32 passes in one launch over a concatenated contraction axis):
            fp32-grade products, used to pin the forward AND backward formulation against the fp32 reference
            (BASELINE.json north star: "TF32 path <= 1e-3"; tests/test_precision_gpu.py).

The mode is read when a forward pass starts and stored with the saved state, so a backward pass always runs in the
mode of its forward. The two fp32 modes run the same kernel templates as the product path with the storage type
swapped (see include/terragan_b200.h, "fp32-storage verification path"); they are not tuned for speed.
"""
from __future__ import annotations

from contextlib import contextmanager

import torch

MODES = ("bf16", "tf32", "tf32x3")
_mode = "bf16"


def set_precision(mode: str) -> None:
    global _mode
    if mode not in MODES:
        raise ValueError(f"precision must be one of {MODES}, got {mode!r}")
    _mode = mode


def get_precision() -> str:
    return _mode


@contextmanager
def precision(mode: str):
    prev = get_precision()
    set_precision(mode)
    try:
        yield
    finally:
        set_precision(prev)


def act_dtype(mode: str) -> torch.dtype:
    return torch.bfloat16 if mode == "bf16" else torch.float32
