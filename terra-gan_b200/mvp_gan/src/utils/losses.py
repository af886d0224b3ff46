"""Drop-in InpaintingLoss / HumanGuidedLoss / BoundaryAwareLoss for the B200 path — same class names,
constructor signatures, attributes (`.boundary_weight`, `.boundary_loss`, `.vgg_layers`, `.l1_loss`,
`.total_variation_loss`, overridden `.to`) and forward contracts as the reference
mvp_gan/src/utils/losses.py:10-428.

What changes underneath: L1 + TV + boundary loss are ONE fused reduction kernel (and one gradient
kernel) with no host synchronisation; the perceptual term runs VGG16 features[:16] as tcgen05
implicit-GEMM convolutions on bf16 channels-last activations with the three replicated input channels
folded into one. `vgg_layers` remains a real nn.Sequential holding the (frozen) fp32 weights.

VGG weights: like the reference, `vgg16(weights=IMAGENET1K_V1)` is tried first. With no network the
download fails; set TERRA_VGG_SEED=<int> (or pass vgg_state_dict=...) to use seeded random weights —
which is what the benchmark spec asks for ("random-init weights"). Without either, the constructor
raises exactly as the reference does.
"""
from typing import Dict, Optional, Tuple
import logging
import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from tg_b200.functional import InpaintTermsFn, PerceptualFn
from tg_b200.layers import VGG_CONVS, VggEngine

logger = logging.getLogger(__name__)


def _build_vgg_features(vgg_state_dict=None) -> nn.Sequential:
    """torchvision vgg16().features[:16] (reference losses.py:31-32)."""
    from torchvision.models import vgg16, VGG16_Weights
    if vgg_state_dict is not None:
        feats = vgg16(weights=None).features[:16]
        feats.load_state_dict(vgg_state_dict)
        return feats
    seed = os.environ.get("TERRA_VGG_SEED")
    if seed is not None:
        import math
        gen = torch.Generator().manual_seed(int(seed))
        feats = vgg16(weights=None).features[:16]
        with torch.no_grad():
            for idx, cin, cout in VGG_CONVS:
                feats[idx].weight.copy_(torch.randn((cout, cin, 3, 3), generator=gen) * math.sqrt(2.0 / (cin * 9)))
                feats[idx].bias.copy_(0.05 * torch.randn((cout,), generator=gen))
        return feats
    return vgg16(weights=VGG16_Weights.IMAGENET1K_V1).features[:16]


class BoundaryAwareLoss(nn.Module):
    """Boundary-weighted L1 (reference losses.py:206-428; only `forward`, :386-428, is on the path —
    the Sobel / gradient-consistency helpers of the reference are never called by it)."""

    def __init__(self, boundary_width: int = 10, epsilon: float = 1e-6, device=None):
        super().__init__()
        self.boundary_width = boundary_width
        self.epsilon = epsilon
        if device is None:
            self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        else:
            self.device = device

    def forward(self, pred: torch.Tensor, target: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
        pred, target, mask = pred.to(self.device), target.to(self.device), mask.to(self.device)
        # empty boundary -> 0, NaN/Inf -> 0 are handled on the device (reference :411, :419 sync the host)
        return InpaintTermsFn.apply(pred, target, mask, 2, float(self.epsilon))[2]


class InpaintingLoss(nn.Module):
    def __init__(self,
                 perceptual_weight: float = 0.1,
                 tv_weight: float = 0.1,
                 boundary_weight: float = 0.5,
                 device=None,
                 vgg_state_dict=None):
        super().__init__()
        if device is None:
            self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        else:
            self.device = device
        self.l1_loss = nn.L1Loss().to(self.device)
        self.perceptual_weight = perceptual_weight
        self.tv_weight = tv_weight
        self.boundary_weight = boundary_weight
        logger.info(f"Initializing VGG model on device: {self.device}")
        try:
            self.vgg_layers = _build_vgg_features(vgg_state_dict).eval().to(self.device)
            for param in self.vgg_layers.parameters():
                param.requires_grad = False
        except Exception as e:
            logger.error(f"Error initializing VGG model: {str(e)}")
            raise
        self.boundary_loss = BoundaryAwareLoss(device=self.device)
        self.boundary_loss = self.boundary_loss.to(self.device)
        logger.info("InpaintingLoss initialized successfully")

    @property
    def _vgg_engine(self):
        e = self.__dict__.get("_vgg_engine_cache")
        if e is None:
            e = VggEngine()
            self.__dict__["_vgg_engine_cache"] = e
        return e

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_vgg_engine_cache", None)
        return d

    def to(self, device):
        """Override to() to ensure all internal components are moved to the device (reference :45-56)."""
        self.device = device
        self.l1_loss = self.l1_loss.to(device)
        self.vgg_layers = self.vgg_layers.to(device)
        if hasattr(self, 'boundary_loss'):
            self.boundary_loss = self.boundary_loss.to(device)
            self.boundary_loss.device = device
        return super().to(device)

    def _vgg_params(self) -> Dict[str, torch.Tensor]:
        out = {}
        for idx, _, _ in VGG_CONVS:
            out[f"{idx}.weight"] = self.vgg_layers[idx].weight
            out[f"{idx}.bias"] = self.vgg_layers[idx].bias
        return out

    def forward(self, input, target, mask):
        input = input.to(self.device)
        target = target.to(self.device)
        mask = mask.to(self.device)
        # L1 (:73), TV on the inpainted region (:96-100) and boundary loss (:106-110): one fused pass
        flags = 0 if self.tv_weight > 0 else 2
        terms = InpaintTermsFn.apply(input, target, mask, flags, float(self.boundary_loss.epsilon))
        total_loss = terms[0]
        if self.perceptual_weight > 0:                                               # :77-90
            perceptual_loss = PerceptualFn.apply(input, target, self._vgg_engine, self._vgg_params())
            total_loss = total_loss + self.perceptual_weight * perceptual_loss
        if self.tv_weight > 0:
            total_loss = total_loss + self.tv_weight * terms[1]
        if self.boundary_weight > 0:
            total_loss = total_loss + self.boundary_weight * terms[2]
        return total_loss

    def total_variation_loss(self, x):
        """TV of an arbitrary tensor (reference :118-127) — mask of ones-complement = no masking."""
        b, c, h, w = x.shape
        # the fused kernel works on single-channel tiles: fold C into the batch axis. The reference divides by
        # batch_size = x.size(0) once more (:127), so the folded result (divided by b*c) is scaled back by c.
        xf = x.reshape(b * c, 1, h, w)
        zeros = torch.zeros_like(xf)
        tv = InpaintTermsFn.apply(xf, xf.detach(), zeros, 0, 1e-6)[1]
        return tv * float(c) if c != 1 else tv

    def _tensor_size(self, t):
        return t.numel()


class HumanGuidedLoss(InpaintingLoss):
    def __init__(self, config, device=None, **kwargs):
        if 'device' in kwargs:
            del kwargs['device']
        boundary_weight = config['training'].get('loss_weights', {}).get('boundary', 0.5)
        super().__init__(device=device, boundary_weight=boundary_weight, **kwargs)
        self.human_feedback_weight = config['training']['modes']['human_guided']['human_feedback_weight']
        self.base_loss_weight = config['training']['modes']['human_guided']['base_loss_weight']
        logger.info(f"HumanGuidedLoss initialized with weights: base={self.base_loss_weight}, "
                    f"human={self.human_feedback_weight}, boundary={boundary_weight}")

    def forward(self, input, target, mask, human_feedback=None):
        input = input.to(self.device)
        target = target.to(self.device)
        mask = mask.to(self.device)
        base_loss = super().forward(input, target, mask)
        human_loss = torch.zeros((), device=input.device)
        if human_feedback is not None and 'mask' in human_feedback and human_feedback['mask'] is not None:
            human_mask = human_feedback['mask'].to(self.device)
            human_guided_regions = (human_mask > 0).float()
            # reference :171 branches on `.sum() > 0` (host sync); with an all-zero mask both terms below
            # are exactly 0 on the device, so the result is identical without the sync.
            terms = InpaintTermsFn.apply(input, target, human_guided_regions, 3, float(self.boundary_loss.epsilon))
            human_loss = terms[0]                                                    # :172-175
            if self.boundary_weight > 0:                                             # :178-185
                human_loss = human_loss + self.boundary_weight * terms[2]
        total_loss = (self.base_loss_weight * base_loss + self.human_feedback_weight * human_loss)   # :197-200
        return total_loss
