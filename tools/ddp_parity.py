"""Multi-GPU numerical parity of the data-parallel train step (SURVEY.md §8e contract):

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/ddp_parity.py [--mode tf32x3|bf16]

Every rank runs tg_b200.step.AdversarialStep on its own shard of a global batch with the bucketed NCCL
reducer attached. After `reducer.finish()` each parameter's .grad must equal the MEAN over ranks of the
per-shard gradients — "reference per shard + gradient average", local BatchNorm statistics and loss normalisers
(DDP semantics). The per-shard reference is the oracle (fp32, and for bf16 the rounding-emulating oracle) run on
the same shard with the rank's own branch decisions replayed (tests/gates.py); rank 0 gathers them, averages and
compares. Also checks that all replicas hold identical parameters after the optimizer steps. Prints one JSON line.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "terra-gan_b200"), os.path.join(ROOT, "tests")]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", default="tf32x3", choices=["bf16", "tf32x3"])
    ap.add_argument("--tile", type=int, default=256)
    ap.add_argument("--batch", type=int, default=2, help="tiles per rank")
    ap.add_argument("--bucket-bytes", type=int, default=4 << 20)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)

    from oracle import terra_oracle as O
    from tg_b200 import precision as PR
    from tg_b200.ddp import BucketedGradReducer, broadcast_module_state
    from tg_b200.step import AdversarialStep
    from mvp_gan.src.models.generator import PConvUNet
    from mvp_gan.src.models.discriminator import Discriminator
    from mvp_gan.src.utils.losses import InpaintingLoss
    import gates as GT

    H, B = args.tile, args.batch
    real = O.make_tiles(100 + rank, B, H)                       # a different shard per rank
    masks = O.make_mask(200 + rank, B, H, "large" if rank % 2 else "rect")
    vgg = O.make_vgg_state(3)
    G, D = PConvUNet(), Discriminator()
    if rank == 0:                                               # other ranks start from garbage: broadcast must fix it
        G.load_state_dict(O.make_generator_state(1))
        D.load_state_dict(O.make_discriminator_state(2))
    G.to(dev).train()
    D.to(dev).train()
    broadcast_module_state([G, D])
    crit = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=dev, vgg_state_dict=vgg)
    reducer = BucketedGradReducer([G, D], bucket_bytes=args.bucket_bytes)
    opt_G, opt_D = torch.optim.Adam(G.parameters(), lr=2e-4), torch.optim.Adam(D.parameters(), lr=2e-4)
    # the reference graph (D weight gradients of the G step computed, then zeroed) so that D's reducer traffic in the
    # G step is exercised too
    stepper = AdversarialStep(G, D, crit, opt_G, opt_D, reducer, skip_discarded_d_wgrad=True)

    snap = {}
    orig_g, orig_d = opt_G.step, opt_D.step

    def step_g(*a, **k):                                        # .grad after reducer.finish(), before the update
        snap["g"] = {n: p.grad.detach().clone() for n, p in G.named_parameters() if p.grad is not None}
        return orig_g(*a, **k)

    def step_d(*a, **k):
        snap["d"] = {n: p.grad.detach().clone() for n, p in D.named_parameters() if p.grad is not None}
        return orig_d(*a, **k)

    opt_G.step, opt_D.step = step_g, step_d
    GT.arm(G, D, crit)
    with PR.precision(args.mode):
        out = stepper.run(real.to(dev), masks.to(dev))
    torch.cuda.synchronize()
    gates = GT.collect(G, D, crit)
    gates["sign.pixel"] = torch.sign(out["gen_imgs"].cpu() - real)
    GT.disarm(G, D, crit)

    # per-shard reference on this rank's CPU (own branch decisions replayed)
    import contextlib
    rounding = O.rounding(O.bf16_ste) if args.mode == "bf16" else contextlib.nullcontext()
    with rounding, O.gate_tape(gates):
        r = O.adversarial_step(real, masks, O.make_generator_state(1), O.make_discriminator_state(2), vgg)
    ref = {"g": r["g_grads"], "d": r["d_grads"]}

    # gather the per-shard references on rank 0 and average them (gloo group for CPU tensors)
    cpu_group = dist.new_group(backend="gloo")
    rows, worst = [], 0.0
    for which in ("g", "d"):
        for name in sorted(ref[which]):
            t = ref[which][name].clone().contiguous()
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=cpu_group)
            t /= world
            got = snap[which][name].float().cpu()
            scale = t.abs().max().item()
            comp = None
            if name.endswith("input_conv.bias"):
                comp = name.replace("input_conv.bias", "bn.bias")
            for ci, bi in ((2, 3), (5, 6), (8, 9)):
                if name == f"model.{ci}.bias":
                    comp = f"model.{bi}.bias"
            if comp is not None:
                c = ref[which][comp].clone().contiguous()
                dist.all_reduce(c, op=dist.ReduceOp.SUM, group=cpu_group)
                scale = max(scale, (c / world).abs().max().item())
            e = ((got - t).abs().max() / max(scale, 1e-30)).item()
            rows.append((which + ":" + name, e))
    rows.sort(key=lambda x: -x[1])
    # every rank must hold the same averaged gradient and, after Adam, the same parameters
    sync_err = 0.0
    for m in (G, D):
        for p in m.parameters():
            a = p.detach().clone()
            dist.broadcast(a, src=0)
            sync_err = max(sync_err, (a - p.detach()).abs().max().item())
    t = torch.tensor([sync_err], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tol = 1e-3 if args.mode == "tf32x3" else 6e-2
    if rank == 0:
        print(json.dumps({"test": "ddp_gradient_parity", "world": world, "mode": args.mode, "tile": H, "batch_per_rank": B,
                          "buckets_launched": reducer.buckets_launched, "tensors": len(rows), "tol": tol,
                          "worst": [(n, float(f"{e:.3e}")) for n, e in rows[:6]],
                          "replica_param_max_abs_diff_after_step": float(t.item()),
                          "ok": bool(rows[0][1] < tol and t.item() == 0.0)}), flush=True)
    ok = rows[0][1] < tol and t.item() == 0.0
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
