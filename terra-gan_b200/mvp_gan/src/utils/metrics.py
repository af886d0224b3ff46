"""Drop-in for the device-side part of the reference's mvp_gan/src/utils/metrics.py — `PerformanceMetrics`
(PSNR / SSIM / L1 / L2, metrics.py:10-46) and `TrainingMetrics` (gradient norms, learning rates, :48-69) — as used
by ExperimentTracker.log_training_batch (utils/experiment_tracking.py:678-695) every `log_interval` batches of the
train loops (train.py:229-266). Same class / method names, argument meaning and return types.

The reference evaluates PSNR, SSIM and L1/L2 with ~35 ATen kernels and five `.item()` host syncs per logged batch;
here they come out of ONE fused reduction (tg_quality_metrics, csrc/metrics_kernels.cu) whose nine results stay on
the device until somebody asks for them: `quality_metrics(pred, target, mask)` returns a lazy `QualityMetrics`, the
static methods below read single fields from it (one D2H copy of 36 bytes, cached per (pred, target) pair so the
three reference calls psnr / ssim / l1_l2 on the same tensors cost one kernel). The host-side resource probes of the
reference file (psutil / GPUtil, `ResourceMetrics`, `MetricsLogger`) are out of scope.
"""
from typing import Dict, Optional, Tuple

import torch

from tg_b200 import ops


class QualityMetrics:
    """The nine numbers of tg_quality_metrics, on the device (`.values`) until `.to_dict()` / `[name]` is used."""

    def __init__(self, values: torch.Tensor):
        self.values = values
        self._host = None

    def _get(self):
        if self._host is None:
            self._host = [float(v) for v in self.values.tolist()]      # the only host synchronisation
        return self._host

    def __getitem__(self, name: str) -> float:
        return self._get()[ops.QUALITY_FIELDS.index(name)]

    def to_dict(self) -> Dict[str, float]:
        return dict(zip(ops.QUALITY_FIELDS, self._get()))


def quality_metrics(pred: torch.Tensor, target: torch.Tensor, mask: Optional[torch.Tensor] = None) -> QualityMetrics:
    """PSNR, SSIM, L1, L2, MSE and (with `mask`) the boundary-quality metrics of [B,1,H,W] tensors in one launch."""
    if not pred.is_cuda:
        raise RuntimeError(f"quality_metrics: input is on {pred.device}; the B200 TERRA-GAN path runs hand-written "
                           "sm_100a CUDA kernels only and has no CPU fallback")
    f = lambda t: t.detach().float().contiguous()
    return QualityMetrics(ops.quality_metrics(f(pred), f(target), f(mask) if mask is not None else None))


_last = {"key": None, "value": None}


def _cached(pred: torch.Tensor, target: torch.Tensor) -> QualityMetrics:
    key = (pred.data_ptr(), pred._version, target.data_ptr(), target._version, tuple(pred.shape))
    if _last["key"] != key:
        _last["key"], _last["value"] = key, quality_metrics(pred, target)
    return _last["value"]


class PerformanceMetrics:
    @staticmethod
    def calculate_psnr(pred: torch.Tensor, target: torch.Tensor) -> float:
        """Peak Signal-to-Noise Ratio, max pixel 1.0 (metrics.py:12-19; inf when pred == target)."""
        return _cached(pred, target)["psnr"]

    @staticmethod
    def calculate_ssim(pred: torch.Tensor, target: torch.Tensor, window_size: int = 11) -> float:
        """Mean SSIM with an 11x11 box window (metrics.py:22-39)."""
        if window_size != 11:
            raise NotImplementedError("calculate_ssim (B200 path): window_size 11 only (the reference default, the only "
                                      "value its callers use: experiment_tracking.py:211, evaluation/metrics.py:57)")
        return _cached(pred, target)["ssim"]

    @staticmethod
    def calculate_l1_l2(pred: torch.Tensor, target: torch.Tensor) -> Tuple[float, float]:
        """(mean |p - t|, sqrt(mean (p - t)^2)) (metrics.py:42-46)."""
        m = _cached(pred, target)
        return m["l1_distance"], m["l2_distance"]


class TrainingMetrics:
    @staticmethod
    def calculate_gradient_norm(model: torch.nn.Module) -> Dict[str, float]:
        """Per-parameter and total gradient L2 norms (metrics.py:50-64): one multi-tensor norm kernel and ONE
        device->host copy instead of an `.item()` per parameter (74 for the generator)."""
        names, grads = [], []
        for name, p in model.named_parameters():
            if p.grad is not None:
                names.append(name)
                grads.append(p.grad.data)
        if not grads:
            return {"total_grad_norm": 0.0}
        norms = torch.stack(torch._foreach_norm(grads, 2)).tolist()
        out = {f"grad_norm_{n}": float(v) for n, v in zip(names, norms)}
        out["total_grad_norm"] = float(sum(v * v for v in norms) ** 0.5)
        return out

    @staticmethod
    def get_learning_rates(optimizer: torch.optim.Optimizer) -> Dict[str, float]:
        return {f"lr_group_{i}": group['lr'] for i, group in enumerate(optimizer.param_groups)}
