"""Drop-in PConvUNet for the B200 path — same no-arg constructor, children (enc1..enc7, dec7..dec1,
final), 114-entry state_dict and forward(x, mask) -> Tensor[B,1,H,W] as the reference
mvp_gan/src/models/generator.py:8-84.

forward() runs the whole U-Net as ONE autograd node (tg_b200.functional.GeneratorFn): mask pyramid
in integers, tcgen05 implicit-GEMM convs with fused PConv epilogues, fused BN/activation and
upsample-concat kernels, bf16 channels-last activations; backward is hand-scheduled the same way.
`decode_step` / `_pad_to_match` are kept for API compatibility (they run the stand-alone layers).
"""
import torch
import torch.nn as nn
from torch.nn.functional import interpolate

from tg_b200.functional import GeneratorFn
from tg_b200.layers import DEC, ENC, BNParams, GeneratorEngine
from .pconv import PConv2d


class PConvUNet(nn.Module):
    def __init__(self):
        super(PConvUNet, self).__init__()
        # Encoder layers — reference generator.py:13-19
        self.enc1 = PConv2d(1, 64, kernel_size=7, stride=2, padding=3)
        self.enc2 = PConv2d(64, 128, kernel_size=5, stride=2, padding=2)
        self.enc3 = PConv2d(128, 256, kernel_size=5, stride=2, padding=2)
        self.enc4 = PConv2d(256, 512, kernel_size=3, stride=2, padding=1)
        self.enc5 = PConv2d(512, 512, kernel_size=3, stride=2, padding=1)
        self.enc6 = PConv2d(512, 512, kernel_size=3, stride=2, padding=1)
        self.enc7 = PConv2d(512, 512, kernel_size=3, stride=2, padding=1)
        # Decoder layers — reference generator.py:22-29
        self.dec7 = PConv2d(512 + 512, 512, kernel_size=3, padding=1)
        self.dec6 = PConv2d(512 + 512, 512, kernel_size=3, padding=1)
        self.dec5 = PConv2d(512 + 512, 512, kernel_size=3, padding=1)
        self.dec4 = PConv2d(512 + 256, 256, kernel_size=3, padding=1)
        self.dec3 = PConv2d(256 + 128, 128, kernel_size=3, padding=1)
        self.dec2 = PConv2d(128 + 64, 64, kernel_size=3, padding=1)
        self.dec1 = PConv2d(64, 64, kernel_size=3, padding=1)
        self.final = nn.Conv2d(64, 1, kernel_size=3, padding=1)

    # ---- engine plumbing (derived state; never pickled) ----
    @property
    def _engine(self):
        e = self.__dict__.get("_engine_cache")
        if e is None:
            e = GeneratorEngine()
            self.__dict__["_engine_cache"] = e
        return e

    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_engine_cache", None)
        d.pop("_trace", None)
        return d

    def _param_items(self):
        names, tensors = [], []
        for name, *_ in ENC + DEC:
            layer = getattr(self, name)
            for sub, attr in (("input_conv", "weight"), ("input_conv", "bias"), ("bn", "weight"), ("bn", "bias")):
                names.append(f"{name}.{sub}.{attr}")
                tensors.append(getattr(getattr(layer, sub), attr))
        names += ["final.weight", "final.bias"]
        tensors += [self.final.weight, self.final.bias]
        return names, tensors

    def _bn_params(self):
        out = {}
        for name, *_ in ENC + DEC:
            bn = getattr(self, name).bn
            out[name] = BNParams(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked)
        return out

    def forward(self, x, mask):
        names, tensors = self._param_items()
        if not torch.is_grad_enabled():
            # inference (evaluate.py:47-50 runs under no_grad): nothing is saved for backward, and in eval mode
            # the decoder BatchNorm + ReLU fold into the conv epilogues (tg_b200.layers.GeneratorEngine.forward)
            if not x.is_cuda:
                raise RuntimeError(f"PConvUNet: input is on {x.device}; the B200 TERRA-GAN path runs hand-written "
                                   "sm_100a CUDA kernels only and has no CPU fallback")
            return self._engine.forward(x, mask, dict(zip(names, tensors)), self._bn_params(), self.training, None,
                                        getattr(self, "_trace", None))
        return GeneratorFn.apply(x, mask, self, tuple(names), *tensors)

    # ---- reference helper API (generator.py:66-84), expressed with the stand-alone layers ----
    def decode_step(self, up_feature, up_mask, skip_feature, skip_mask, decoder_layer):
        up_feature = interpolate(up_feature, scale_factor=2, mode='bilinear', align_corners=False)
        up_mask = interpolate(up_mask, scale_factor=2, mode='nearest')
        up_feature = self._pad_to_match(up_feature, skip_feature)
        up_mask = self._pad_to_match(up_mask, skip_mask)
        merged_feature = torch.cat([up_feature, skip_feature], dim=1)
        merged_mask = torch.max(up_mask, skip_mask)
        out_feature, out_mask = decoder_layer(merged_feature, merged_mask)
        return out_feature, out_mask

    def _pad_to_match(self, x, target):
        """Pads tensor x to match the size of target tensor along spatial dimensions."""
        diffY = target.size(2) - x.size(2)
        diffX = target.size(3) - x.size(3)
        x = nn.functional.pad(x, [diffX // 2, diffX - diffX // 2,
                                  diffY // 2, diffY - diffY // 2])
        return x
