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
# train: 690.163 GF per tile as the reference graph computes it; 13.035 GF of that is the D weight gradient of the
# generator step, which train.py:210 zeroes unused and this path skips by default (--ref-graph computes it): 677.128
GFLOP_PER_TILE = {"train": 690.163, "train_skip_d_wgrad": 677.128, "infer": 93.634, "hg": 573.118}
PCONV_GFLOP_PER_TILE = {"train": 280.490, "hg": 280.490, "infer": 93.634}   # the 14 PConv layers + final conv
TILE = 512
METRIC = "gan_train_step_dsm_tiles_per_sec"
METRICS = {"train": METRIC, "infer": "generator_inference_dsm_tiles_per_sec", "hg": "hg_finetune_step_dsm_tiles_per_sec"}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1383.8), d.get("hbm_gbs", 6551.4), "measured", d.get("bf16_tflops", 1632.5)
    return 1400.0, 6650.0, "fallback", 1650.0


def make_masks(kind: str, B: int, gen: "torch.Generator"):
    """Synthetic hole masks (1 = valid), SURVEY.md §8d. rect: 1-4 axis-aligned holes of side 32-256 px (M_rect).
    large: union of big rectangles and blobs to 50-80 % hole so that holes survive to enc6 / enc7 (M_large, the
    "large-mask irregular holes at high mask density" of BASELINE.json configs[4])."""
    m = torch.ones(B, 1, TILE, TILE)
    if kind == "rect":
        for b in range(B):
            for _ in range(int(torch.randint(1, 5, (1,), generator=gen))):
                hh, ww = (int(torch.randint(32, 257, (1,), generator=gen)) for _ in range(2))
                y0 = int(torch.randint(0, TILE - hh + 1, (1,), generator=gen))
                x0 = int(torch.randint(0, TILE - ww + 1, (1,), generator=gen))
                m[b, 0, y0:y0 + hh, x0:x0 + ww] = 0
        return m
    for b in range(B):
        target = 0.5 + 0.3 * float(torch.rand((1,), generator=gen))
        # irregular blobs: threshold a smooth random field (low-resolution noise, bilinearly up-sampled) ...
        field = torch.nn.functional.interpolate(torch.rand((1, 1, 16, 16), generator=gen), size=(TILE, TILE),
                                                mode="bicubic", align_corners=False)[0, 0]
        thr = torch.quantile(field.flatten(), target * 0.8)
        m[b, 0][field < thr] = 0
        # ... plus rectangles until the target hole density is reached
        while float(1 - m[b].mean()) < target:
            hh, ww = (int(torch.randint(64, 257, (1,), generator=gen)) for _ in range(2))
            y0 = int(torch.randint(0, TILE - hh + 1, (1,), generator=gen))
            x0 = int(torch.randint(0, TILE - ww + 1, (1,), generator=gen))
            m[b, 0, y0:y0 + hh, x0:x0 + ww] = 0
    return m


def stock_torch_gpu_baseline(dev, batch: int, steps: int):
    """Informational bar (SURVEY.md §2.2 / BASELINE.md §4): the reference's own ATen ops on this GPU, i.e. what
    stock PyTorch + cuDNN does with the unmodified algorithm — the oracle port with every tensor on the device, in
    fp32 and with TF32 convolutions allowed. (/root/reference does not exist on the GPU box; the port is verified
    bit-identical to the reference modules on CPU by tests/test_oracle_golden.py.)"""
    from oracle import terra_oracle as O
    to = lambda sd: {k: v.to(dev) for k, v in sd.items()}
    out = {}
    real, masks = O.make_tiles(5, batch, TILE).to(dev), O.make_mask(6, batch, TILE, "rect").to(dev)
    for name, tf32 in (("fp32", False), ("tf32", True)):
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = tf32
        g_sd, d_sd, vgg, opt = to(O.make_generator_state(1)), to(O.make_discriminator_state(2)), to(O.make_vgg_state(3)), {}
        try:
            for _ in range(2):
                O.adversarial_step(real, masks, g_sd, d_sd, vgg, lr=2e-4, opt_state=opt)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                O.adversarial_step(real, masks, g_sd, d_sd, vgg, lr=2e-4, opt_state=opt)
            e1.record()
            torch.cuda.synchronize()
            out[name] = batch / (e0.elapsed_time(e1) / steps * 1e-3)
        except Exception as exc:          # never let the informational leg take the bench line down
            out[name] = f"failed: {type(exc).__name__}: {exc}"[:200]
        del g_sd, d_sd, vgg, opt
        torch.cuda.empty_cache()
    torch.backends.cudnn.allow_tf32 = True
    return {"unit": "tiles/s", "batch": batch, "steps": steps, "tiles_per_s": out,
            "what": "oracle port of train.py:179-225 (the reference's own ATen ops, torch.autograd, cuDNN/cuBLAS) on this "
                    "GPU; informational, not the reference arm"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index: int):
        self.rows, self.proc, self.index, self.t_mark = [], None, index, 0.0

    def mark(self):
        """Start of the timed region: only samples taken from here on are reported (the sampler itself is started before
        the warm-up steps, because nvidia-smi needs up to a second to deliver its first sample)."""
        self.t_mark = time.monotonic()

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        inside = [r for t, r in self.rows if t >= self.t_mark]
        note = None
        if not inside and self.rows:       # timed region shorter than the sampling period: the last sample of the warm-up (same load)
            inside, note = [self.rows[-1][1]], "timed region shorter than the 100 ms sampling period: last warm-up sample"
        for r in inside:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        out = {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
               "samples": len(sm)}
        if note:
            out["note"] = note
        return out


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

    if args.global_batch:
        if args.global_batch % world:
            raise RuntimeError(f"--global-batch {args.global_batch} is not divisible by {world} ranks")
        B = args.global_batch // world          # strong scaling (configs[4]) / fixed global batch (configs[3])
    else:
        B = args.batch
    scaling = "strong" if args.global_batch else "weak"
    mask_kind = args.masks or ("large" if args.workload == "hg" else "rect")
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
    mask_h = make_masks(mask_kind, B, gen).pin_memory()
    human_h = (1 - make_masks("rect", B, gen)).pin_memory()      # human-flagged regions (1 = flagged), configs[4]
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

    human_d = human_h.to(dev)
    graphed = None
    if args.workload == "infer" and not args.no_graph:
        from tg_b200.graphs import GraphedGenerator
        graphed = GraphedGenerator(G, B, TILE, TILE)     # one cudaGraphLaunch per forward (small batches are launch-bound)

    def run_step(r, m, h=None):
        if args.workload == "infer":
            if graphed is not None:
                out = graphed(r * m, m)
            else:
                with torch.no_grad():
                    out = G(r * m, m)
            return {"g_total_loss": out.mean(), "d_loss": out.mean(), "out": out}
        if args.workload == "hg":
            o = hg_stepper.run(r, m, human_d if h is None else h)
            return {"g_total_loss": o["loss"], "d_loss": o["loss"]}
        return stepper.run(r, m)

    def step_resident():
        run_step(real_d, mask_d)

    # ---- end to end: every step's inputs come from pinned host memory. Like a DataLoader(pin_memory=True) feeding
    # `.to(device, non_blocking=True)`, the copy of step i+1 is issued on a copy stream while step i computes; the
    # compute stream waits for its own batch's copy event. Results are read back every step: the two loss scalars
    # (train.py:222-225) for the training workloads, the inpainted tiles (evaluate.py:53) for inference.
    copy_stream = torch.cuda.Stream()
    n_in = 3 if args.workload == "hg" else 2
    host_in = [real_h, mask_h, human_h][:n_in]
    bufs = [[torch.empty_like(t, device=dev) for t in host_in] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    state = {"i": 0}
    loss_host = torch.empty(2, pin_memory=True)
    out_host = torch.empty((B, 1, TILE, TILE), pin_memory=True) if args.workload == "infer" else None

    def prefetch(slot):
        copy_stream.wait_event(consumed[slot])           # the previous user of this slot is done with it
        with torch.cuda.stream(copy_stream):
            for dst, src in zip(bufs[slot], host_in):
                dst.copy_(src, non_blocking=True)
            ready[slot].record()

    def step_e2e():
        slot = state["i"] & 1
        if state["i"] == 0:
            prefetch(0)
        prefetch(slot ^ 1)                               # next step's inputs travel while this step computes
        torch.cuda.current_stream().wait_event(ready[slot])
        out = run_step(*bufs[slot])
        consumed[slot].record()
        if out_host is not None:
            out_host.copy_(out["out"], non_blocking=True)
        else:
            loss_host.copy_(torch.stack([out["g_total_loss"], out["d_loss"]]), non_blocking=True)
        torch.cuda.current_stream().synchronize()        # the caller reads the result of every step
        state["i"] += 1

    ops.profile_pool(2 * 130 * args.steps + 64)      # timing events for every tensor-core launch of the timed steps
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step_resident()
    # ---- timed region 1: resident inputs; tensor-core launches timed with CUDA events ----
    sampler.mark()
    _lib.CALLS.clear()
    for o in (opt_G, opt_D, getattr(hg_stepper, "opt", None)):
        if hasattr(o, "kernel_launches"):
            o.kernel_launches = 0
    ops.PROFILE = []
    if graphed is not None:
        ms_step = timed(step_resident, args.steps)      # replayed graph: launches are not individually timed
        prof = []
        ops.PROFILE = None
        with torch.no_grad():                            # per-launch profile + launch count from one eager forward each
            ops.PROFILE = []
            _lib.CALLS.clear()
            for _ in range(args.steps):
                G(real_d * mask_d, mask_d)
            torch.cuda.synchronize()
            prof, ops.PROFILE = ops.PROFILE, None
    else:
        ms_step = timed(step_resident, args.steps)
        prof, ops.PROFILE = ops.PROFILE, None
    launches = _lib.kernel_launches()
    fused = [o for o in (opt_G, opt_D, getattr(hg_stepper, "opt", None)) if hasattr(o, "kernel_launches")]
    if fused and args.workload != "infer":      # the fused Adam launches one kernel per 24 tensors, not one per call
        launches += sum(o.kernel_launches for o in fused) - _lib.CALLS.get("tg_adam_repack", 0)
    calls = dict(_lib.CALLS)
    # ---- timed region 2: end to end from host buffers ----
    step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    clocks = sampler.stop()

    tc = {}
    if args.per_launch and rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", args.per_launch), "w") as f:
            for kind, flops, a, b, shape, tag in prof[: len(prof) // max(args.steps, 1)]:
                ms = a.elapsed_time(b)
                f.write(f"{tag:3s} {kind:10s} {ms*1e3:9.1f} us {flops/ms/1e9:8.1f} TFLOP/s  {flops/1e9:9.1f} GF  {shape}\n")
    pconv = {}                                           # the generator's 14 PConv layers + final conv, incl. enc1 / final
    for kind, flops, a, b, _shape, tag in prof:
        ms = a.elapsed_time(b)
        if tag == "G":
            t = pconv.setdefault(kind, [0.0, 0.0, 0])
            t[0] += flops
            t[1] += ms
            t[2] += 1
        if kind.startswith("thin_"):
            continue                                     # bandwidth-bound 1<->64-channel convs: not part of the GEMM family
        t = tc.setdefault(kind, [0.0, 0.0, 0])
        t[0] += flops
        t[1] += ms
        t[2] += 1
    tc_flops = sum(v[0] for v in tc.values())
    tc_ms = sum(v[1] for v in tc.values())
    peak_tf, peak_bw, peak_src, peak_burst = load_peaks()
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    pc_flops = sum(v[0] for v in pconv.values())
    pc_ms = sum(v[1] for v in pconv.values())
    pc_tf = pc_flops / (pc_ms * 1e-3) / 1e12 if pc_ms > 0 else 0.0
    pc_gemm = {k: v for k, v in pconv.items() if not k.startswith("thin_")}
    pcg_flops, pcg_ms = sum(v[0] for v in pc_gemm.values()), sum(v[1] for v in pc_gemm.values())
    pcg_tf = pcg_flops / (pcg_ms * 1e-3) / 1e12 if pcg_ms > 0 else 0.0

    # DRAM traffic of the same kernel family from the committed ncu --set full capture (one B=64 train step)
    traffic, traffic_note = None, None
    for tname in ("r02_ncu_tensorcore_step_summary.json", "r01_ncu_tensorcore_step_summary.json"):
        tpath = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tpath) and B == 64 and args.workload == "train":
            t = json.load(open(tpath))
            traffic = t["dram_gbytes"] * 1e9 / t["launches"]
            traffic_note = (f"dram__bytes_read+write: {t['dram_gbytes']:.1f} GB over the {t['launches']} tensor-core launches of "
                            f"one B=64 step (ncu --set full capture summarised in profiles/{tname}); bytes per launch (mean)")
            break
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    tiles_per_step = B * world
    value = tiles_per_step / (ms_step * 1e-3)
    e2e = tiles_per_step / (ms_e2e * 1e-3)
    cpu = None
    stock = None
    if world == 1 and not args.no_cpu_baseline:
        if args.workload == "train":
            stock = stock_torch_gpu_baseline(dev, 16, 3)
        dt, threads = cpu_reference_step_time(4, 4, 1)
        cpu = {"value": 4 / dt, "unit": "tiles/s", "cores": threads, "kind": "port",
               "sample": "1 warm-up + 4 timed adversarial steps at batch 4 (oracle/terra_oracle.py: the reference's "
                         "own ATen/oneDNN ops, fp32, all host threads)"}
    wl_key = "train_skip_d_wgrad" if (args.workload == "train" and not args.ref_graph) else args.workload
    step_gf = GFLOP_PER_TILE[wl_key]
    line = {
        "metric": METRICS[args.workload], "value": value, "unit": "tiles/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": {"train": "adversarial train step (train.py:179-225): PConvUNet + Discriminator + "
                                         "InpaintingLoss (perceptual 0.1, tv 0.1, boundary 0.5) + Adam x2, 1x512x512 DSM "
                                         "tiles",
                                "infer": "generator inference, eval mode, no_grad (evaluate.py:47-50), 1x512x512 DSM tiles"
                                         + ("" if args.no_graph else ", one CUDA-graph launch per forward"),
                                "hg": "human-guided fine-tune step (human_guided_trainer.py:101-153): PConvUNet + "
                                      "HumanGuidedLoss (0.7/0.3, boundary 0.5, human masks) + Adam 1e-4, 1x512x512 DSM "
                                      "tiles"}[args.workload],
                   "masks": {"rect": "M_rect: 1-4 rectangular holes of side 32-256 px",
                             "large": "M_large: irregular blobs + rectangles, 50-80 % hole"}[mask_kind],
                   "tile": TILE, "batch_per_gpu": B, "global_batch": tiles_per_step,
                   "parallelism": f"dp{world}", "l2": "working set per step (>10 GB) exceeds the 126 MB L2",
                   "d_wgrad_in_g_step": "computed (reference graph)" if args.ref_graph else
                                        "skipped (zeroed unused by train.py:210; output-equivalent)",
                   "step_gflop_per_tile": step_gf,
                   "step_frac_of_bf16_peak": value / world * step_gf * 1e9 / (peak_tf * 1e12)},
        "e2e": {"value": e2e, "unit": "tiles/s", "ms_per_step": ms_e2e,
                "h2d_bytes_per_step": int(sum(t.numel() * 4 for t in host_in)) * world,
                "d2h_bytes_per_step": (int(out_host.numel() * 4) if out_host is not None else 8) * world,
                "input_pipeline": "double-buffered: step i+1's pinned-host -> device copy runs on a copy stream during step i"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel + wgrad_igemm_kernel (tcgen05 implicit GEMM)",
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "peak_source": f"{peak_src} bf16_tflops_sustained", "traffic": traffic,
                     "traffic_note": traffic_note,
                     "share_of_step": tc_ms / (ms_step * args.steps) if ms_step > 0 else None,
                     "by_kind": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0,
                                     "ms_per_step": v[1] / args.steps, "launches_per_step": v[2] / args.steps}
                                 for k, v in tc.items()},
                     # BASELINE.json's second half: "PConv tensor-pipe % of peak" = the generator's PConv layers alone
                     # (enc1..enc7, dec7..dec1 and the final conv; fprop + dgrad + wgrad launches issued by the
                     # generator engine), algorithmic FLOPs / summed CUDA-event time of those launches
                     "pconv_only": {
                         "tflops": pc_tf, "ms_per_step": pc_ms / args.steps,
                         "gflop_per_tile": pc_flops / args.steps / B / 1e9,
                         "frac_of_sustained_peak": pc_tf / peak_tf, "frac_of_burst_peak": pc_tf / peak_burst,
                         "tensor_core_layers": {"tflops": pcg_tf, "ms_per_step": pcg_ms / args.steps,
                                                "frac_of_sustained_peak": pcg_tf / peak_tf,
                                                "frac_of_burst_peak": pcg_tf / peak_burst,
                                                "note": "enc2..enc7, dec7..dec1 (implicit GEMM); enc1 and the final conv are "
                                                        "bandwidth-bound 1<->64-channel kernels, included in the line above"},
                         "peaks": {"sustained": peak_tf, "burst": peak_burst},
                         "by_kind": {k: {"tflops": v[0] / (v[1] * 1e-3) / 1e12 if v[1] > 0 else 0.0,
                                         "ms_per_step": v[1] / args.steps} for k, v in pconv.items()}}},
        "cpu_baseline": cpu,
        "stock_pytorch_on_this_gpu": stock,
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
    ap.add_argument("--global-batch", type=int, default=0,
                    help="fix the GLOBAL batch (tiles per step over all ranks): strong scaling (configs[4]: 256) or the "
                         "fixed 512 of configs[3]; per-GPU batch = global / ranks. Default: --batch per GPU (weak scaling)")
    ap.add_argument("--masks", default="", choices=["", "rect", "large"], help="hole-mask family (default: rect; large for hg)")
    ap.add_argument("--no-graph", action="store_true", help="infer workload: eager launches instead of one CUDA graph")
    ap.add_argument("--per-launch", default="", help="write a per-launch table of the tensor-core kernels to gpurun_out/<name>")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
