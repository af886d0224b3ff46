#!/usr/bin/env python
"""bench.py — TERRA-GAN hot-path benchmark on B200 (contract: see README / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl ours|reference]

Metric (BASELINE.json): GAN train-step DSM tiles/sec. One "step" = one full adversarial train step
(mvp_gan/src/train.py:179-225: G fwd, InpaintingLoss incl. VGG perceptual / TV / boundary 0.5, D fwd,
BCE, G bwd, Adam(G), D fwd x2, BCE, D bwd, Adam(D)) on one batch of synthetic 1x512x512 DSM tiles and
random hole masks, random-init weights (BASELINE.json configs[2]: batch 64 per B200; weak scaling:
64 tiles per GPU). Prints ONE JSON line on rank 0.

  value     tiles/s, inputs already resident in HBM, CUDA-event timed, max over ranks
  e2e       same metric through the public module API with HOST inputs: per step H2D copy of the
            tiles + masks from pinned memory and D2H read of the two loss scalars inside the timed region
  roofline  the tensor-core implicit-GEMM kernels (fprop+dgrad+wgrad): algorithmic FLOPs / CUDA-event
            time of those launches inside the timed steps, against the measured cuBLAS bf16 peak
  cpu_baseline  the oracle port of the reference (oracle/terra_oracle.py) timed on this host's cores

--impl reference times that CPU path as its own arm (the reference ships no GPU kernels of its own
and cannot be pip-installed: it is a script tree, not a package).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "terra-gan_b200"))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

GFLOP_PER_TILE_STEP = 690.163          # algorithmic work of one adversarial step per tile
GFLOP_PER_TILE = {"train": 690.163, "infer": 93.634, "hg": 573.118}
TILE = 512
METRIC = "gan_train_step_dsm_tiles_per_sec"
METRICS = {"train": METRIC, "infer": "generator_inference_dsm_tiles_per_sec", "hg": "hg_finetune_step_dsm_tiles_per_sec"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1383.8), d.get("hbm_gbs", 6551.4), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the oracle port of the reference's train step
# ------------------------------------------------------------------------------------------------
def cpu_reference_step_time(batch: int, steps: int, warmup: int):
    from oracle import terra_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    g_sd, d_sd, vgg = O.make_generator_state(1), O.make_discriminator_state(2), O.make_vgg_state(3)
    real = O.make_tiles(5, batch, TILE)
    masks = O.make_mask(6, batch, TILE, "rect")
    opt = {}
    for _ in range(warmup):
        O.adversarial_step(real, masks, g_sd, d_sd, vgg, lr=2e-4, opt_state=opt)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.adversarial_step(real, masks, g_sd, d_sd, vgg, lr=2e-4, opt_state=opt)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 2
    dt, threads = cpu_reference_step_time(batch, args.steps, args.warmup)
    tps = batch / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": tps, "unit": "tiles/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "adversarial train step (train.py:179-225), 1x512x512 tiles, CPU sample batch 2",
                   "tile": TILE, "batch_per_step": batch},
        "cpu_baseline": {"value": tps, "unit": "tiles/s", "cores": threads, "kind": "port",
                         "sample": f"{args.steps} timed + {args.warmup} warm-up adversarial steps at batch {batch} "
                                   "(oracle/terra_oracle.py restating train.py:179-225 with the reference's own ATen ops)"},
        "e2e": {"value": tps, "unit": "tiles/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py: no CUDA device — the TERRA-GAN B200 path has no CPU fallback "
                           "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from tg_b200 import _lib, ops
    from tg_b200.ddp import BucketedGradReducer, broadcast_module_state
    from tg_b200.step import AdversarialStep
    from mvp_gan.src.models.generator import PConvUNet
    from mvp_gan.src.models.discriminator import Discriminator
    from mvp_gan.src.utils.losses import InpaintingLoss

    B = args.batch
    torch.manual_seed(1)                               # random-init weights (default PyTorch init), same on every rank
    G, D = PConvUNet(), Discriminator()
    G.to(dev).train()
    D.to(dev).train()
    os.environ.setdefault("TERRA_VGG_SEED", "3")       # no network: seeded random VGG16[:16] weights
    criterion = InpaintingLoss(perceptual_weight=0.1, tv_weight=0.1, device=dev)
    if args.torch_adam:                                # the reference loops' optimizer object, unchanged
        make_adam = lambda m, lr: torch.optim.Adam(m.parameters(), lr=lr)
    else:                                              # same update, one fused kernel that also refreshes the packed weights
        from tg_b200 import optim as tg_optim
        make_adam = lambda m, lr: tg_optim.Adam(m.parameters(), lr=lr, modules=[m])
    opt_G = make_adam(G, 2e-4)
    opt_D = make_adam(D, 2e-4)
    reducer = None
    if world > 1:
        broadcast_module_state([G, D])
        reducer = BucketedGradReducer([G, D])
    stepper = AdversarialStep(G, D, criterion, opt_G, opt_D, reducer, skip_discarded_d_wgrad=not args.ref_graph)
    hg_stepper = None
    if args.workload == "hg":          # BASELINE.json configs[4]: human-guided fine-tuning step (no discriminator)
        from tg_b200.step import HumanGuidedStep
        from mvp_gan.src.utils.losses import HumanGuidedLoss
        cfg = {"training": {"loss_weights": {"perceptual": 0.1, "tv": 0.1, "boundary": 0.5},
                            "modes": {"human_guided": {"human_feedback_weight": 0.3, "base_loss_weight": 0.7,
                                                       "learning_rate": 1e-4}}}}
        hg_crit = HumanGuidedLoss(cfg, device=dev)
        hg_stepper = HumanGuidedStep(G, hg_crit, make_adam(G, 1e-4), reducer)
    if args.workload == "infer":       # BASELINE.json configs[1]: generator inference (evaluate.py:47-50)
        G.eval()

    # synthetic DSM tiles + rectangular hole masks, a different shard per rank; pinned host copies for e2e
    gen = torch.Generator().manual_seed(1234 + rank)
    real_h = torch.rand((B, 1, TILE, TILE), generator=gen).pin_memory()
    mask_h = torch.ones(B, 1, TILE, TILE)
    for b in range(B):
        for _ in range(int(torch.randint(1, 5, (1,), generator=gen))):
            hh, ww = (int(torch.randint(32, 257, (1,), generator=gen)) for _ in range(2))
            y0 = int(torch.randint(0, TILE - hh + 1, (1,), generator=gen))
            x0 = int(torch.randint(0, TILE - ww + 1, (1,), generator=gen))
            mask_h[b, 0, y0:y0 + hh, x0:x0 + ww] = 0
    mask_h = mask_h.pin_memory()
    real_d, mask_d = real_h.to(dev), mask_h.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / steps

    human_d = (torch.rand((B, 1, TILE, TILE), generator=gen) < 0.1).float().to(dev)

    def run_step(r, m):
        if args.workload == "infer":
            with torch.no_grad():
                out = G(r * m, m)
            return {"g_total_loss": out.mean(), "d_loss": out.mean()}
        if args.workload == "hg":
            o = hg_stepper.run(r, m, human_d)
            return {"g_total_loss": o["loss"], "d_loss": o["loss"]}
        return stepper.run(r, m)

    def step_resident():
        run_step(real_d, mask_d)

    loss_host = torch.empty(2, pin_memory=True)

    def step_e2e():
        r = real_h.to(dev, non_blocking=True)
        m = mask_h.to(dev, non_blocking=True)
        out = run_step(r, m)
        loss_host.copy_(torch.stack([out["g_total_loss"], out["d_loss"]]), non_blocking=True)
        torch.cuda.current_stream().synchronize()     # the user reads the losses every step (train.py:222-225)

    ops.profile_pool(2 * 100 * args.steps + 64)      # timing events for every tensor-core launch of the timed steps
    for _ in range(args.warmup):
        step_resident()
    # ---- timed region 1: resident inputs; tensor-core launches timed with CUDA events ----
    sampler = ClockSampler(local)
    sampler.start()
    _lib.CALLS.clear()
    ops.PROFILE = []
    ms_step = timed(step_resident, args.steps)
    prof, ops.PROFILE = ops.PROFILE, None
    launches = _lib.kernel_launches()
    calls = dict(_lib.CALLS)
    # ---- timed region 2: end to end from host buffers ----
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()

    tc = {}
    if args.per_launch and rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", args.per_launch), "w") as f:
            for kind, flops, a, b, shape in prof[: len(prof) // max(args.steps, 1)]:
                ms = a.elapsed_time(b)
                f.write(f"{kind:6s} {ms*1e3:9.1f} us {flops/ms/1e9:8.1f} TFLOP/s  {flops/1e9:9.1f} GF  {shape}\n")
    for kind, flops, a, b, _shape in prof:
        t = tc.setdefault(kind, [0.0, 0.0, 0])
        t[0] += flops
        t[1] += a.elapsed_time(b)
        t[2] += 1
    tc_flops = sum(v[0] for v in tc.values())
    tc_ms = sum(v[1] for v in tc.values())
    peak_tf, peak_bw, peak_src = load_peaks()
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0

    # DRAM traffic of the same kernel family from the committed ncu --set full capture (one B=64 train step)
    traffic, traffic_note = None, None
    tpath = os.path.join(ROOT, "profiles", "r01_ncu_tensorcore_step_summary.json")
    if os.path.exists(tpath) and B == 64 and args.workload == "train":
        t = json.load(open(tpath))
        traffic = t["dram_gbytes"] * 1e9 / t["launches"]
        traffic_note = (f"dram__bytes_read+write: {t['dram_gbytes']:.1f} GB over the {t['launches']} tensor-core launches of one "
                        f"B=64 step (profiles/r01_ncu_tensorcore_step_metrics.csv); bytes per launch (mean)")
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tiles_per_step = B * world
    value = tiles_per_step / (ms_step * 1e-3)
    e2e = tiles_per_step / (ms_e2e * 1e-3)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        dt, threads = cpu_reference_step_time(4, 4, 1)
        cpu = {"value": 4 / dt, "unit": "tiles/s", "cores": threads, "kind": "port",
               "sample": "1 warm-up + 4 timed adversarial steps at batch 4 (oracle/terra_oracle.py: the reference's "
                         "own ATen/oneDNN ops, fp32, all host threads)"}
    line = {
        "metric": METRICS[args.workload], "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": {"train": "adversarial train step (train.py:179-225): PConvUNet + Discriminator + "
                                         "InpaintingLoss (perceptual 0.1, tv 0.1, boundary 0.5) + Adam x2, 1x512x512 DSM "
                                         "tiles, rect hole masks",
                                "infer": "generator inference, eval mode, no_grad (evaluate.py:47-50), 1x512x512 DSM tiles",
                                "hg": "human-guided fine-tune step (human_guided_trainer.py:101-153): PConvUNet + "
                                      "HumanGuidedLoss (0.7/0.3, boundary 0.5) + Adam 1e-4, 1x512x512 DSM tiles"}[args.workload],
                   "tile": TILE, "batch_per_gpu": B, "global_batch": tiles_per_step,
                   "parallelism": f"dp{world}", "l2": "working set per step (>10 GB) exceeds the 126 MB L2",
                   "d_wgrad_in_g_step": "computed (reference graph)" if args.ref_graph else
                                        "skipped (zeroed unused by train.py:210; output-equivalent)",
                   "step_gflop_per_tile": GFLOP_PER_TILE[args.workload],
                   "step_frac_of_bf16_peak": value / world * GFLOP_PER_TILE[args.workload] * 1e9 / (peak_tf * 1e12)},
        "e2e": {"value": e2e, "unit": "tiles/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(real_h.numel() * 4 + mask_h.numel() * 4) * world,
                "d2h_bytes_per_step": 8 * world},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel + wgrad_igemm_kernel (tcgen05 implicit GEMM)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": traffic,
                     "traffic_note": traffic_note,
                     "share_of_step": tc_ms / (ms_step * args.steps) if ms_step > 0 else None,
                     "by_kind": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0,
                                     "ms_per_step": v[1] / args.steps, "launches_per_step": v[2] / args.steps}
                                 for k, v in tc.items()}},
        "cpu_baseline": cpu,
        "clocks": clocks,
        "calls_per_step": {k: v / args.steps for k, v in sorted(calls.items())},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=64, help="tiles per GPU per step (BASELINE.json configs[2]: 64)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-graph", action="store_true",
                    help="also compute the discriminator weight gradients of the generator step (discarded by the reference)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--torch-adam", action="store_true", help="use torch.optim.Adam instead of tg_b200.optim.Adam")
    ap.add_argument("--workload", default="train", choices=["train", "infer", "hg"],
                    help="train: adversarial step (headline, configs[2]); infer: generator inference (configs[1]); "
                         "hg: human-guided fine-tune step (configs[4])")
    ap.add_argument("--per-launch", default="", help="write a per-launch table of the tensor-core kernels to gpurun_out/<name>")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
