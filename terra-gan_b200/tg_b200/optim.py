"""Fused Adam for the B200 path: one multi-tensor kernel per optimizer step that applies
torch.optim.Adam's update (amsgrad off, weight_decay 0 — what mvp_gan/src/train.py:61-62 and
training/human_guided_trainer.py:63 construct) and, for the convolutions that run on the tensor cores, writes the
packed bf16 fprop / dgrad operand matrices in the same pass, so the engines never re-pack after a step.

    opt_G = tg_b200.optim.Adam(generator.parameters(), lr=2e-4, modules=[generator])

is a drop-in for `torch.optim.Adam(generator.parameters(), lr=2e-4)`: same constructor arguments, same
`state_dict()` layout (`step`, `exp_avg`, `exp_avg_sq` per parameter), `zero_grad`, `param_groups`.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, Optional

import torch

from . import plan as P
from ._lib import AdamTensor, check, lib, ptr, stream_ptr


def _conv_packs(module) -> Dict[int, object]:
    """id(weight parameter) -> ConvPack of the engine of a drop-in PConvUNet / Discriminator."""
    eng = module._engine
    out = {}
    if hasattr(eng, "packs"):          # GeneratorEngine: every PConv layer but enc1 (1 input channel, thin kernel)
        for name, pk in eng.packs.items():
            w = getattr(module, name).input_conv.weight
            if w.shape[1] % 64 == 0:
                out[id(w)] = pk
    if hasattr(eng, "_packs"):         # DiscriminatorEngine: model[2], model[5], model[8]
        for idx, pk in eng._packs.items():
            out[id(module.model[idx].weight)] = pk
    return out


class Adam(torch.optim.Optimizer):
    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 amsgrad: bool = False, modules: Optional[Iterable] = None):
        if weight_decay != 0.0 or amsgrad:
            raise ValueError("tg_b200.optim.Adam implements the configuration the reference uses: weight_decay=0, amsgrad=False")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0.0, amsgrad=False))
        self._packs: Dict[int, object] = {}
        for m in (modules or []):
            self._packs.update(_conv_packs(m))
        self._index: Dict[tuple, tuple] = {}    # (id(param), device) -> (dst_fprop, dst_dgrad) int32 scatter indices
        self.kernel_launches = 0                # kernels launched by step() so far (bench.py reads and resets it)

    def attach(self, module) -> None:
        """Also refresh the packed weights of this drop-in module's tensor-core convolutions."""
        self._packs.update(_conv_packs(module))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            by_step, touched = {}, []      # one launch per distinct step count (a parameter whose grad was None on some
            #                                steps, or a loaded state_dict with differing steps, lags behind the others)
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("tg_b200.optim.Adam: parameters must be contiguous fp32 CUDA tensors")
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                s = int(st["step"].item())          # host tensor: no device synchronisation
                e = AdamTensor()
                e.param, e.grad, e.exp_avg, e.exp_avg_sq = ptr(p), ptr(g), ptr(st["exp_avg"]), ptr(st["exp_avg_sq"])
                e.n = p.numel()
                pk = self._packs.get(id(p))
                if pk is not None:
                    wf, wd = pk.w_fprop(p), pk.w_dgrad(p)       # builds the packed copies on first use
                    ikey = (id(p), str(p.device))              # the module may have been moved since the last step
                    idx = self._index.get(ikey)
                    if idx is None:
                        idx = (P.pack_scatter_index(p.shape, None, p.device), P.pack_scatter_index(p.shape, pk.dplan, p.device))
                        self._index[ikey] = idx
                    e.packed_fprop, e.dst_fprop, e.packed_dgrad, e.dst_dgrad = ptr(wf), ptr(idx[0]), ptr(wd), ptr(idx[1])
                    touched.append((pk, p))
                by_step.setdefault(s, []).append((e, g))
            if not by_step:
                continue
            for step_no, entries in sorted(by_step.items()):
                arr = (AdamTensor * len(entries))(*[e for e, _ in entries])
                check(lib().tg_adam_repack(arr, len(entries), float(group["lr"]), float(beta1), float(beta2),
                                           float(group["eps"]), step_no, stream_ptr()), "tg_adam_repack")
                self.kernel_launches += (len(entries) + 23) // 24       # one kernel per 24 tensors (bench.py counts launches)
            for pk, p in touched:
                torch.autograd.graph.increment_version(p)       # the master changed behind autograd's back
                pk.mark_fresh(p)                                # ... and its packed copies are already current
            for p in group["params"]:
                if p.grad is not None and id(p) not in self._packs:
                    torch.autograd.graph.increment_version(p)
        return loss
