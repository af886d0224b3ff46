"""Drop-in Discriminator for the B200 path — same constructor, `.model` nn.Sequential with the
parametrised indices 0,2,3,5,6,8,9,11 (25-entry state_dict) and forward(img) -> patch logits as the
reference mvp_gan/src/models/discriminator.py:6-26.

The Sequential holds the fp32 master parameters only; forward() runs one autograd node
(tg_b200.functional.DiscriminatorFn): 4x4/s2 convs as tcgen05 implicit GEMMs on parity-split bf16
activations, BN statistics from the conv epilogue, fused BN + LeakyReLU, bandwidth kernels for the
1-channel first and last convolutions.
"""
import torch
import torch.nn as nn

from tg_b200.functional import DiscriminatorFn
from tg_b200.layers import DISC_MID, BNParams, DiscriminatorEngine


class Discriminator(nn.Module):
    def __init__(self, input_channels=1):
        super(Discriminator, self).__init__()
        if input_channels != 1:
            raise NotImplementedError("Discriminator (B200 path): input_channels=1 only (the reference default, "
                                      "the only value its loops use: train.py:107, main_pipeline.py:214)")

        def discriminator_block(in_channels, out_channels, normalization=True):
            layers = [nn.Conv2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1)]
            if normalization:
                layers.append(nn.BatchNorm2d(out_channels))
            layers.append(nn.LeakyReLU(0.2, inplace=True))
            return layers

        self.model = nn.Sequential(
            *discriminator_block(input_channels, 64, normalization=False),
            *discriminator_block(64, 128),
            *discriminator_block(128, 256),
            *discriminator_block(256, 512),
            nn.Conv2d(512, 1, kernel_size=4, padding=1)
        )

    @property
    def _engine(self):
        e = self.__dict__.get("_engine_cache")
        if e is None:
            e = DiscriminatorEngine()
            self.__dict__["_engine_cache"] = e
        return e

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_engine_cache", None)
        return d

    def _param_items(self):
        names, tensors = [], []
        for idx in (0, 2, 3, 5, 6, 8, 9, 11):
            for attr in ("weight", "bias"):
                names.append(f"model.{idx}.{attr}")
                tensors.append(getattr(self.model[idx], attr))
        return names, tensors

    def _bn_params(self):
        out = {}
        for _, bi, _, _ in DISC_MID:
            bn = self.model[bi]
            out[bi] = BNParams(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked)
        return out

    def forward(self, img):
        names, tensors = self._param_items()
        if not torch.is_grad_enabled():      # nothing to save for backward
            if not img.is_cuda:
                raise RuntimeError(f"Discriminator: input is on {img.device}; the B200 TERRA-GAN path runs hand-written "
                                   "sm_100a CUDA kernels only and has no CPU fallback")
            return self._engine.forward(img, dict(zip(names, tensors)), self._bn_params(), self.training, None)
        return DiscriminatorFn.apply(img, self, tuple(names), *tensors)
