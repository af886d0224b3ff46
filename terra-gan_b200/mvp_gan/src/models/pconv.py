"""Drop-in PConv2d for the B200 path — same constructor, attributes, state_dict keys and
forward(input, mask) -> (output, output_mask) contract as the reference
mvp_gan/src/models/pconv.py:6-50, computed by hand-written sm_100a kernels (tg_b200).

Parameters stay ordinary fp32 nn.Conv2d / nn.BatchNorm2d sub-modules (`input_conv`, `mask_conv`,
`bn`) so checkpoints load in both directions; the sub-modules are parameter containers only — their
own forward is never called. There is no CPU fallback: inputs must live on a CUDA device.
"""
import torch
import torch.nn as nn

from tg_b200.functional import PConv2dFn
from tg_b200.layers import ConvPack


def _as_int(v):
    return v[0] if isinstance(v, (tuple, list)) else v


class PConv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, batch_norm=True):
        super(PConv2d, self).__init__()
        self.input_conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride, padding, bias=True)
        self.slide_winsize = self.input_conv.weight.data.shape[2] * self.input_conv.weight.data.shape[3]
        self.mask_conv = nn.Conv2d(1, 1, kernel_size, stride, padding, bias=False)
        torch.nn.init.constant_(self.mask_conv.weight, 1.0)       # reference pconv.py:14
        for param in self.mask_conv.parameters():
            param.requires_grad = False
        self.batch_norm = batch_norm
        if self.batch_norm:
            self.bn = nn.BatchNorm2d(out_channels)
        self.activation = nn.ReLU()
        self._k, self._stride, self._pad = _as_int(kernel_size), _as_int(stride), _as_int(padding)
        if self.input_conv.kernel_size[0] != self.input_conv.kernel_size[1]:
            raise NotImplementedError("PConv2d (B200 path): square kernels only (all reference layers are square)")

    @property
    def _pack(self):
        pk = self.__dict__.get("_pack_cache")
        if pk is None:
            pk = ConvPack(self._k, self._stride, self._pad)
            self.__dict__["_pack_cache"] = pk
        return pk

    def __getstate__(self):            # derived device caches are not part of the pickled module
        d = dict(self.__dict__)
        d.pop("_pack_cache", None)
        return d

    def forward(self, input, mask):
        gamma = self.bn.weight if self.batch_norm else None
        beta = self.bn.bias if self.batch_norm else None
        return PConv2dFn.apply(input, mask, self, self.input_conv.weight, self.input_conv.bias, gamma, beta)
