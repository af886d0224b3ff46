"""Drop-in for the device-side function of the reference's mvp_gan/src/evaluation/metrics.py:
`calculate_boundary_quality(pred, target, mask, boundary_width=10)` (:79-133), which train.py:237-246 and
human_guided_trainer.py:128-137 call every `log_interval` batches, plus the PSNR / SSIM / MSE bundle of
`MaskEvaluator.calculate_metrics` (:135-149) as `calculate_metrics`. The OpenCV contour / IoU part of `MaskEvaluator`
(:22-45) is host-side analysis and out of scope.

All numbers come from the fused reduction tg_quality_metrics (csrc/metrics_kernels.cu): one launch, one 36-byte
device->host copy — the reference uses ~15 ATen kernels and four host syncs for the boundary metrics alone.
"""
from typing import Dict

import torch

from ..utils.metrics import quality_metrics


def calculate_boundary_quality(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor,
                               boundary_width: int = 10) -> Dict[str, float]:
    """{'boundary_mse', 'boundary_psnr', 'boundary_gradient_diff'} exactly as the reference defines them
    (`boundary_width` is accepted and, as in the reference, unused: the boundary is the 3x3 dilate - erode band)."""
    m = quality_metrics(pred, target, mask)
    return {k: m[k] for k in ("boundary_mse", "boundary_psnr", "boundary_gradient_diff")}


def calculate_metrics(pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> Dict[str, float]:
    """mse, psnr, ssim + the boundary metrics (MaskEvaluator.calculate_metrics, :135-149) from one launch."""
    m = quality_metrics(pred, target, mask)
    return {k: m[k] for k in ("mse", "psnr", "ssim", "boundary_mse", "boundary_psnr", "boundary_gradient_diff")}
