"""Data-parallel gradient exchange for the B200 path: one process per GPU, replicated weights, the
global batch of tiles sharded over ranks, ONE exchange step per optimiser step — a bucketed
sum-allreduce (NCCL over NVLink 5 / NVSwitch) of the gradients, launched on a side stream from inside
the hand-scheduled backward as soon as each layer's gradients are final, so it overlaps the rest of
backward. The reference has no distributed code at all (single device, train.py:50);
semantics follow PyTorch DDP: local BatchNorm statistics, local loss normalisers, gradient average.

The engines call `on_grads(module, names, tensors)` (tg_b200.functional.set_grad_hook) in
reverse-forward order: final, dec1..dec7, enc7..enc1 for the generator; model.11 .. model.0 for the
discriminator. Tensors are packed into ~25 MB buckets; each bucket is flattened, all-reduced and
written into `param.grad` on the communication stream by `finish()`, which must be called after
backward() and before optimizer.step(): it flushes the last bucket, replaces every reduced parameter's
.grad by the rank average and makes the compute stream wait for the exchange. (The flat bucket is a
copy taken when the layer's gradient is final, so it does not matter whether autograd later keeps or
clones the tensor it was handed; a parameter that receives several contributions in one backward —
the discriminator is applied to two batches in the D step — is summed, which commutes with averaging.)

Works with any torch.distributed backend (tests run it on CPU with gloo, world_size 2).
"""
from __future__ import annotations

from typing import List, Optional

import torch
import torch.distributed as dist

from . import functional


class BucketedGradReducer:
    def __init__(self, modules, bucket_bytes: int = 25 * 1024 * 1024, process_group=None):
        self.modules = list(modules)
        self.bucket_bytes = bucket_bytes
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self._pending: List[torch.Tensor] = []
        self._pending_keys: List[tuple] = []
        self._pending_bytes = 0
        self._inflight = []         # (work, flat bucket, [(module, param name)], [numel], source tensors)
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self.buckets_launched = 0
        for m in self.modules:
            functional.set_grad_hook(m, self._on_grads)

    def close(self) -> None:
        for m in self.modules:
            functional.set_grad_hook(m, None)

    # ---- called from inside backward ----
    def _on_grads(self, module, names, tensors) -> None:
        if self.world == 1:
            return
        for n, t in zip(names, tensors):
            if t is None:
                continue
            self._pending.append(t)
            self._pending_keys.append((module, n))
            self._pending_bytes += t.numel() * t.element_size()
        if self._pending_bytes >= self.bucket_bytes:
            self._flush()

    def _flush(self) -> None:
        if not self._pending:
            return
        tensors, self._pending, self._pending_bytes = self._pending, [], 0
        keys, self._pending_keys = self._pending_keys, []
        cuda = tensors[0].is_cuda
        if cuda:
            if self._comm_stream is None:
                # high priority: the compute kernels are persistent (one CTA per SM), so a collective can only start at
                # a kernel boundary; with priority it is scheduled first when SMs free up instead of queueing behind
                # the next convolution
                self._comm_stream = torch.cuda.Stream(priority=-1)
            ready = torch.cuda.Event()
            ready.record()                                  # gradients of this bucket are complete here
            self._comm_stream.wait_event(ready)
            with torch.cuda.stream(self._comm_stream):
                flat = torch.cat([t.reshape(-1) for t in tensors])
                work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        else:
            flat = torch.cat([t.reshape(-1) for t in tensors])
            work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
        # `tensors` were allocated on the compute stream but are read by the cat on the comm stream: keep them alive
        # until finish() so the caching allocator cannot hand their memory to a later backward kernel first
        self._inflight.append((work, flat, keys, [t.numel() for t in tensors], tensors))
        self.buckets_launched += 1

    # ---- called by the training step before optimizer.step() ----
    def finish(self) -> None:
        if self.world == 1:
            return
        self._flush()
        inflight, self._inflight = self._inflight, []
        if not inflight:
            return
        cuda = inflight[0][1].is_cuda
        inv = 1.0 / self.world

        def scatter_back():
            # one multi-tensor copy (and one add for parameters that received several contributions) per step
            # instead of one tiny kernel per parameter: round 1 spent ~78 launches on the comm stream here
            seen = set()
            copy_dst, copy_src, add_dst, add_src = [], [], [], []
            for work, flat, keys, sizes, _src in inflight:
                work.wait()
                flat.mul_(inv)
                off = 0
                for (module, name), n in zip(keys, sizes):
                    p = module.get_parameter(name)
                    if p.grad is None:
                        raise RuntimeError(f"BucketedGradReducer.finish(): {name} has no .grad — call finish() "
                                           "after backward() and before optimizer.step()")
                    src = flat[off:off + n].view_as(p.grad)
                    if (id(module), name) in seen:
                        add_dst.append(p.grad)
                        add_src.append(src)
                    else:
                        copy_dst.append(p.grad)
                        copy_src.append(src)
                        seen.add((id(module), name))
                    off += n
            if copy_dst:
                torch._foreach_copy_(copy_dst, copy_src)
            if add_dst:
                torch._foreach_add_(add_dst, add_src)

        if cuda:
            after_bwd = torch.cuda.Event()
            after_bwd.record()                  # autograd has written every .grad on the compute stream
            self._comm_stream.wait_event(after_bwd)
            with torch.cuda.stream(self._comm_stream):
                scatter_back()
                done = torch.cuda.Event()
                done.record()
            torch.cuda.current_stream().wait_event(done)
        else:
            scatter_back()


def broadcast_module_state(modules, src: int = 0, process_group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (DDP's initial broadcast)."""
    if not dist.is_initialized() or dist.get_world_size(process_group) == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=process_group)
