#!/usr/bin/env python
"""Headline benchmark: rendered rays/s of the generator forward at 128x128, 24+24 hierarchical
samples per ray, batch 8, synthetic 64^3 x 32 feature volume, random-init TALLSIREN_FG
(BASELINE.json configs[1]) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--siren TYPE] [--precision bf16|fp32]

A "step" is one ImplicitGenerator3d.forward over one batch (per GPU: weak scaling; rays shard
across ranks with no data-path collective).  Rank 0 prints ONE JSON line.

  value     rays/s with (volume, global feature, cam2world) already resident in HBM, CUDA-event timed,
            max over ranks
  e2e       rays/s through the public API with HOST buffers: every step copies that step's inputs
            from pinned host memory and reads pixels + depth back (copies inside the timed region)
  roofline  FiLM-SIREN MLP (cng_film_siren_fwd, the dominant kernel): algorithmic FLOPs per launch
            / its CUDA-event duration measured live in a separate instrumented pass, against the
            measured bf16 tensor peak of MEASURED_PEAKS.json
  cpu_baseline  the torch-CPU oracle (a port of the reference's path; kind "port") on a bounded
            sample of the same workload, all host threads
``--impl reference`` times that CPU path alone (rank 0 only) and prints the same line shape.

Other workloads of BASELINE.json (not the driver's default line):
  --workload train   full GAN train step (U-Net encoder + generator + discriminator, config 3) at 128x128, 48+48 samples,
                     GLOBAL batch 32 split over the N ranks; --workload train_generator times the generator's share alone
  (train_generator)  generator-only train step at 128x128, 48+48 samples, GLOBAL batch 32 split over the N
                     ranks (strong scaling): forward with grad + backward kernels + NCCL gradient all-reduce
                     (DDP) + Adam; the encoder and the discriminator are out of scope and absent
  --workload video   256x256, 48+48 samples, 64 poses of one object sharded over the N ranks (config 4)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

METRIC = "rendered_rays_per_sec_128x128_24+24spp"
UNIT = "rays/s"
FOV = 49.134342641202636
WORKLOAD = dict(batch=8, img_size=128, num_steps=24, volume=64, channels=32, z_dim=256)
SIREN_LAYERS = {"TALLSIREN_FG": 8, "SHORTSIREN_FG": 4, "DOUBLESIREN_FG": 2, "SingleSIREN_dg": 1}


def mlp_flops_per_point(L: int, C: int = 32, H: int = 256) -> int:
    """2*(C*256 + (L-1)*256^2 + 256*4)."""
    return 2 * (C * H + (L - 1) * H * H + H * 4)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tflops_burst=p["bf16_tflops"], tflops_sustained=p["bf16_tflops_sustained"], source="measured")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="fallback")


def render_meta(img_size, num_steps):
    # configs/thousand/special.py:35-41 render parameters
    return dict(img_size=img_size, fov=FOV, ray_start=0.25, ray_end=1.95, num_steps=num_steps, hierarchical_sample=True,
                clamp_mode="relu", nerf_noise=0.0, white_back=True)


def synthetic_inputs(batch, volume, seed):
    """Feature volume ~ N(0, 0.3^2), global ~ N(0.19, 0.05^2), look-at cameras on a shell."""
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import create_cam2world_matrix, sample_camera_positions
    g = torch.Generator().manual_seed(seed)
    vol = torch.randn((batch, WORKLOAD["channels"], volume, volume, volume), generator=g) * 0.3
    glob = torch.randn((batch, WORKLOAD["z_dim"]), generator=g) * 0.05 + 0.19
    st = np.random.get_state()
    np.random.seed(seed)                      # sample_camera_positions draws from numpy's global generator, as the reference's does
    cam = create_cam2world_matrix(sample_camera_positions("cpu", "y", 0.7, 1.5, batch), "y", "cpu")
    np.random.set_state(st)
    return vol, glob, cam


def random_init_generator(siren_type):
    """Random-init weights of the named architecture: the module's own constructor applies the reference's init
    (frequency_init / first_layer_film_sine_init, generators/siren.py:40-44, 134-143); seeded so that every rank and both
    bench arms of one run build the same network."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    rng = torch.random.get_rng_state()
    torch.manual_seed(0)
    gen = ImplicitGenerator3d(siren_type, WORKLOAD["z_dim"], WORKLOAD["channels"], 4, 256)
    torch.random.set_rng_state(rng)
    return gen


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path
# --------------------------------------------------------------------------------------------------
def cpu_render_time(siren_type, img_size, num_steps, volume, reps, warmup):
    from oracle import nerf_path as oracle
    torch.set_num_threads(os.cpu_count() or 1)
    state = oracle.init_generator_state(siren_type, seed=0)
    vol, glob, cam = synthetic_inputs(1, volume, 0)
    meta = render_meta(img_size, num_steps)
    g = torch.Generator().manual_seed(1)
    times = []
    for i in range(warmup + reps):
        draws = oracle.draw_randoms(1, img_size, num_steps, True, g)
        t0 = time.perf_counter()
        oracle.render(state, siren_type, (vol, glob), cam, draws, taps=False, **meta)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def pick_cpu_sample(siren_type, budget_s, n_steps):
    """Largest image size in {32, 64, 128} (batch 1, same 24+24 samples, same 64^3 volume) whose n_steps
    renders fit the time budget, from a small calibration render."""
    t = min(cpu_render_time(siren_type, 32, WORKLOAD["num_steps"], WORKLOAD["volume"], 1, 1))
    per_ray = t / (32 * 32)
    for img in (128, 64):
        if per_ray * img * img * n_steps <= budget_s:
            return img
    return 32


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    img = pick_cpu_sample(args.siren, 150.0, args.steps + args.warmup)
    times = cpu_render_time(args.siren, img, WORKLOAD["num_steps"], WORKLOAD["volume"], args.steps, args.warmup)
    total = sum(times)
    value = img * img * args.steps / total
    sample = f"batch 1 of 8, {img}x{img} rays, 24+24 samples, 64^3x32 volume per step (torch-CPU oracle port of the reference path, fp32)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, "cpu"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, where):
    return {"workload": "generator_render_128x128_24+24spp_batch8_vol64^3x32 (BASELINE configs[1])", "siren_type": args.siren,
            "batch_per_gpu": WORKLOAD["batch"], "img_size": WORKLOAD["img_size"], "samples_per_ray": "24+24",
            "feature_volume": "64^3 x 32ch fp32", "precision": args.precision if where == "gpu" else "fp32",
            "l2": "no flush: per-step working set (268 MB volume + 2 x 403 MB gathered features) exceeds the 126 MB L2",
            "parallelism": f"rays/images sharded over {args.gpus} GPU(s), no data-path collective"}


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in open(self.path).read().splitlines():
            f = [x.strip() for x in row.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


def run_gpu(args):
    import torch.distributed as dist

    from conditioned_nerf_gan_b200 import _lib, ops

    _lib.load()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        from conditioned_nerf_gan_b200 import parallel
        parallel.bind_to_gpu_numa_node(local)        # pinned host buffers next to the GPU they feed (end-to-end leg)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    B, img, S, V = WORKLOAD["batch"], WORKLOAD["img_size"], WORKLOAD["num_steps"], WORKLOAD["volume"]
    R = img * img
    L = SIREN_LAYERS[args.siren]
    meta = render_meta(img, S)
    gen = random_init_generator(args.siren)
    gen = gen.to(dev).eval()
    gen.set_device(dev)
    gen.siren.precision = args.precision
    vol_h, glob_h, cam_h = (t.pin_memory() for t in synthetic_inputs(B, V, seed=rank))
    vol, glob, cam = vol_h.to(dev), glob_h.to(dev), cam_h.to(dev)
    # end-to-end leg: the host holds the feature volumes in fp16 (the dtype the encoder emits under the trainer's autocast,
    # utils.py:643-647) unless --e2e-volume fp32; they are widened on the device in the layout pass (cng_volume_f16_to_channels_last)
    vol_e2e_h = vol_h.half().pin_memory() if args.e2e_volume == "fp16" else vol_h
    torch.manual_seed(rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_resident():
        with torch.no_grad():
            return gen((vol, glob), cam, **meta)

    pix_h = torch.empty((B, 3, img, img), dtype=torch.float32).pin_memory()
    dep_h = torch.empty((B, img, img), dtype=torch.float32).pin_memory()

    def step_e2e():
        v, g, c = vol_e2e_h.to(dev, non_blocking=True), glob_h.to(dev, non_blocking=True), cam_h.to(dev, non_blocking=True)
        with torch.no_grad():
            px, dp = gen((v, g), c, **meta)
        pix_h.copy_(px, non_blocking=True)
        dep_h.copy_(dp, non_blocking=True)
        torch.cuda.current_stream().synchronize()      # the caller holds the image on the host

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = ops.launch_count
        start.record()
        for _ in range(steps):
            fn()
        end.record()
        barrier()
        return max_over_ranks(start.elapsed_time(end)), ops.launch_count - n0

    from conditioned_nerf_gan_b200.streaming import render_host_batches

    def e2e_pipelined(steps, volume_h):
        """The public host-buffer API: every step's inputs start in pinned host memory and its image ends there;
        H2D of step i+1 and D2H of step i-1 overlap the kernels of step i (three streams, two buffer sets)."""
        n = 0
        for px_h, dp_h in render_host_batches(gen, ((volume_h, glob_h, cam_h) for _ in range(steps)), meta, device=dev):
            n += 1
        assert n == steps
        torch.cuda.synchronize()

    def timed_e2e(steps, warmup, volume_h=None):
        volume_h = vol_e2e_h if volume_h is None else volume_h
        e2e_pipelined(warmup, volume_h)
        barrier()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        e2e_pipelined(steps, volume_h)
        end.record()
        barrier()
        return max_over_ranks(start.elapsed_time(end))

    sampler = ClockSampler(local) if rank == 0 else None
    ms_total, launches = timed(step_resident, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e_sync, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    # two complete K-step passes; the faster one is reported (host-side copies on shared hosts see neighbour traffic), both are listed
    e2e_passes = [timed_e2e(args.steps, max(3, args.warmup)), timed_e2e(args.steps, 1)]
    ms_e2e = min(e2e_passes)
    # the same with fp32 host volumes (two passes, the faster one: the first re-allocates the device staging buffers)
    ms_e2e_fp32 = min(timed_e2e(args.steps, 3, vol_h), timed_e2e(args.steps, 1, vol_h)) if args.e2e_volume == "fp16" else ms_e2e

    # ---- roofline leg: per-entry-point CUDA-event durations over an instrumented pass of the same steps
    # (the timed steps above go through the one-call cng_render_fwd; here the same kernels are launched entry point by entry
    # point so that each one can be bracketed by events)
    def step_by_stage():
        with torch.no_grad():
            return gen((vol, glob), cam, fused_call=False, **meta)

    step_by_stage()
    ops.kernel_events = {}
    for _ in range(args.steps):
        step_by_stage()
    torch.cuda.synchronize()
    per_kernel = {k: [s.elapsed_time(e) for s, e in v] for k, v in ops.kernel_events.items()}
    ops.kernel_events = None
    mlp_ms = float(np.mean(per_kernel["cng_film_siren_fwd"]))
    step_kernel_ms = {k: float(np.sum(v)) / args.steps for k, v in per_kernel.items()}
    pk = peaks()
    flops_per_launch = mlp_flops_per_point(L) * B * R * S            # one pass (coarse or fine) per launch
    achieved = flops_per_launch / (mlp_ms * 1e-3) / 1e12
    # the kernel is CUDA-event timed inside a sub-second region at ~1.9 GHz: the matching denominator is the BURST bf16 peak
    # (the sustained figure was measured at ~1.3 GHz over seconds; the fraction against it is kept as frac_vs_sustained)
    peak = pk["tflops_burst"] if args.precision != "fp32" else None
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "k2_dram_traffic.json")
    if args.precision != "fp32" and os.path.exists(tpath):
        traffic = json.load(open(tpath))["dram_bytes_per_launch"]      # from the committed ncu --set full capture
    roofline = {"kernel": "film_siren_tc_kernel (cng_film_siren_fwd)" if args.precision != "fp32" else "film_siren_simt_kernel",
                "bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                "frac": (achieved / peak) if peak else None, "traffic": traffic,
                "algorithmic_bytes_per_launch": B * R * S * (WORKLOAD["channels"] * 4 + 16) + B * ((2 + 4 * (L - 1)) * 32768 + 8192),
                "peak_source": f"{pk['source']} bf16 burst (cuBLAS 8192^3 best-of-10; the kernel is event-timed inside a sub-second region)",
                "frac_vs_sustained": (achieved / pk["tflops_sustained"]) if peak else None,
                "traffic_source": "profiles/k2_dram_traffic.json (committed ncu --set full capture of this kernel at this shape, not measured live)" if traffic else None,
                "flops_per_launch": flops_per_launch, "ms_per_launch": mlp_ms,
                "share_of_step": float(np.sum(per_kernel["cng_film_siren_fwd"]) / args.steps / sum(step_kernel_ms.values())),
                "step_ms_by_entry_point": step_kernel_ms}

    # ---- second half of BASELINE's metric: train images/s of config 3 at this N (global batch 32, strong scaling)
    train = None
    if not args.no_train:
        train = measure_train(args, rank, world, local, dev, barrier, max_over_ranks, args.train_steps, 3)
    c5 = eager = None
    if rank == 0 and not args.no_extras:
        c5 = measure_c5()
        eager = measure_gpu_eager(args)
    barrier()
    if rank == 0:
        cores = os.cpu_count() or 1
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            img_cpu = pick_cpu_sample(args.siren, 25.0, 2)
            t = min(cpu_render_time(args.siren, img_cpu, S, V, 1, 1))
            cpu = {"value": img_cpu * img_cpu / t, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"batch 1 of 8, {img_cpu}x{img_cpu} rays, 24+24 samples, 64^3x32 volume, 1 timed render after 1 warm-up "
                             f"(torch-CPU oracle port of the reference path, fp32, {cores} threads)"}
        rays = world * B * R * args.steps
        h2d = vol_e2e_h.numel() * vol_e2e_h.element_size() + glob_h.numel() * 4 + cam_h.numel() * 4
        d2h = pix_h.numel() * 4 + dep_h.numel() * 4
        line = {
            "metric": METRIC, "value": rays / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[args.precision], "data": "synthetic",
            "config": workload_config(args, "gpu"),
            "e2e": {"value": rays / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps, "api": "streaming.render_host_batches (H2D / kernels / D2H on three streams, double-buffered)",
                    "unpipelined_value": rays / (ms_e2e_sync * 1e-3), "passes_ms_per_step": [t / args.steps for t in e2e_passes],
                    "host_volume_dtype": args.e2e_volume, "fp32_host_volume_value": rays / (ms_e2e_fp32 * 1e-3),
                    "fp32_host_volume_h2d_bytes_per_step": vol_h.numel() * 4 + glob_h.numel() * 4 + cam_h.numel() * 4},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "train": train, "gpu_eager_baseline": eager, "c5": c5,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _setup_ranks(args):
    import torch.distributed as dist
    from conditioned_nerf_gan_b200 import _lib, parallel
    _lib.load()
    rank, world, local = parallel.init_distributed("nccl")
    if world > 1:
        parallel.bind_to_gpu_numa_node(local)
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    return rank, world, local, dev, barrier, max_over_ranks


def _timed_steps(fn, steps, warmup, barrier, max_over_ranks):
    from conditioned_nerf_gan_b200 import ops
    for _ in range(warmup):
        fn()
    barrier()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = ops.launch_count
    start.record()
    for _ in range(steps):
        fn()
    end.record()
    barrier()
    return max_over_ranks(start.elapsed_time(end)), ops.launch_count - n0


def measure_train(args, rank, world, local, dev, barrier, max_over_ranks, steps, warmup):
    """Full GAN train step of BASELINE config 3: 3D U-Net encoder + FiLM-SIREN generator (CUDA rendering path, forward and
    backward) + progressive discriminator with R1, following utils.py:621-842 (conditioned_nerf_gan_b200/training.py).
    Global batch 32 split over the ranks (strong scaling); returns the metrics of the timed steps (max over ranks)."""
    import torch.distributed as dist
    from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    from conditioned_nerf_gan_b200.generators.unet3d import UNet3D
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import create_cam2world_matrix, sample_camera_positions
    from conditioned_nerf_gan_b200.training import GanTrainStep
    GLOBAL_B, img, S, V = args.global_batch, 128, 48, 64
    if GLOBAL_B % world:
        raise SystemExit(f"global batch {GLOBAL_B} must divide by the number of GPUs")
    b = GLOBAL_B // world
    rng = torch.random.get_rng_state()
    torch.manual_seed(0)                                            # same random-init weights on every rank (reference init distributions)
    np.random.seed(rank)
    gen = ImplicitGenerator3d(args.siren, 256, 32, 4, 256)
    gen.siren.precision = args.precision
    enc = UNet3D(in_channels=4, out_channels=32, f_maps=32, num_levels=4, is_segmentation=False, final_sigmoid=False, return_global=True)   # configs/thousand/special.py:53-62
    disc = ProgressiveDiscriminator()
    torch.random.set_rng_state(rng)
    gen, enc, disc = gen.to(dev), enc.to(dev), disc.to(dev)
    md = dict(render_meta(img, S), nerf_noise=1.0, batch_split=1, r1_lambda=10, grad_clip=1, betas=(0.0, 0.9), weight_decay=0,
              gen_lr=10e-6, disc_lr=10e-5, enc_lr=2e-5, photo_loss=True, depth_loss=False, depth_loss_weight=1, enable_discriminator=True,
              random_gen_img=True, cam_r_start=0.7, cam_r_end=1.5, fade_steps=2000)                                  # configs/thousand/default.py:42-81, special.py:29-41
    trainer = GanTrainStep(gen, enc, disc, md, dev, amp=True, ddp=world > 1, local_rank=local)
    g = torch.Generator().manual_seed(100 + rank)
    occ = (torch.rand((b, 1, V, V, V), generator=g) < 0.05).float()
    voxel_h = torch.cat([occ, torch.rand((b, 3, V, V, V), generator=g) * occ], dim=1).pin_memory()       # ~5 % occupied, U[0,1) colours
    img_h = (torch.rand((b, 3, img, img), generator=g) * 2 - 1).pin_memory()
    cam_h = create_cam2world_matrix(sample_camera_positions("cpu", "y", 0.7, 1.5, b), "y", "cpu").pin_memory()
    results = []

    def step():
        sample = {"img": img_h.to(dev, non_blocking=True), "voxel": voxel_h.to(dev, non_blocking=True), "cam2world": cam_h.to(dev, non_blocking=True)}
        losses = trainer.step(sample)
        results.append(float(losses["d_loss"].item()) + float(losses["g_loss"].item()))     # device->host read of the step's result

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = _timed_steps(step, steps, warmup, barrier, max_over_ranks)
    clocks = sampler.stop() if sampler else None
    # host time to ISSUE one step (no device wait inside: the losses are not read), against the device time above
    torch.cuda.synchronize()
    t_host = time.perf_counter()
    sample = {"img": img_h.to(dev, non_blocking=True), "voxel": voxel_h.to(dev, non_blocking=True), "cam2world": cam_h.to(dev, non_blocking=True)}
    trainer.step(sample)
    host_issue_ms = (time.perf_counter() - t_host) * 1e3
    torch.cuda.synchronize()
    # the gradient all-reduce alone: one flat buffer of the step's gradient bytes (G + E + D parameters, fp32) over NCCL,
    # timed with CUDA events (inside the step DDP overlaps it with the backward kernels)
    n_param = sum(p.numel() for m in (gen, enc, disc) for p in m.parameters())
    allreduce_ms = 0.0
    if world > 1:
        flat = torch.zeros((n_param,), dtype=torch.float32, device=dev)
        for _ in range(2):
            dist.all_reduce(flat)
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(5):
            dist.all_reduce(flat)
        s1.record()
        barrier()
        allreduce_ms = max_over_ranks(s0.elapsed_time(s1)) / 5
        del flat
    L = SIREN_LAYERS[args.siren]
    pts = GLOBAL_B * img * img * 2 * S
    out = {"images_per_s": GLOBAL_B * steps / (ms * 1e-3), "ms_per_step": ms / steps, "global_batch": GLOBAL_B, "batch_per_gpu": b,
           "steps": steps, "warmup": warmup, "launches": launches // max(steps, 1), "host_issue_ms": host_issue_ms, "allreduce_ms": allreduce_ms,
           "allreduce_bytes": 4 * n_param, "scaling": "strong", "final_loss": results[-1], "clocks": clocks,
           "h2d_bytes_per_step": int((voxel_h.numel() + img_h.numel() + cam_h.numel()) * 4), "d2h_bytes_per_step": 8,
           "mlp_flops_per_step": 3 * 2 * mlp_flops_per_point(L) * pts, "cos_dump_bits": _cos_dump_bits(),
           "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2**30, "reserved_mem_gb": torch.cuda.memory_reserved(dev) / 2**30,
           "workload": "full GAN train step (3D U-Net encoder + FiLM-SIREN generator + progressive discriminator, D step with R1 then G/E step), "
                       "128x128, 48+48 samples/ray, global batch 32 (BASELINE configs[2]); autocast fp16 + GradScaler, Adam x3; "
                       "host voxels / images / cameras copied in and the losses read back every step"}
    del trainer, gen, enc, disc
    torch.cuda.empty_cache()
    return out


def run_train(args):
    import torch.distributed as dist
    rank, world, local, dev, barrier, max_over_ranks = _setup_ranks(args)
    t = measure_train(args, rank, world, local, dev, barrier, max_over_ranks, args.steps, args.warmup)
    if rank == 0:
        line = {"metric": "train_images_per_sec_128x128_48+48spp_global_batch32", "value": t["images_per_s"],
                "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t["ms_per_step"],
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": t["workload"], "siren_type": args.siren, "batch_per_gpu": t["batch_per_gpu"], "batch_split": 1,
                           "amp": "autocast fp16 + GradScaler (as utils.py:643,711)",
                           "optimizers": "Adam x3", "grad_allreduce": "DDP/NCCL, once per optimizer step" if world > 1 else "none",
                           "encoder": "cuDNN (library), channels_last_3d, emits the NDHWC volume zero-copy", "discriminator": "cuDNN (library)",
                           "generator": "hand-written CUDA path: forward x2 (no-grad for the D step, with grad for the G step) + backward"},
                "e2e": {"value": t["images_per_s"], "unit": "images/s", "h2d_bytes_per_step": t["h2d_bytes_per_step"], "d2h_bytes_per_step": 8},
                "gpu_launches": t["launches"] * args.steps, "clocks": t["clocks"], "final_loss": t["final_loss"],
                "allreduce_ms": t["allreduce_ms"], "host_issue_ms": t["host_issue_ms"], "mlp_flops_per_step": t["mlp_flops_per_step"],
                "cos_dump_bits": t.get("cos_dump_bits")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _cos_dump_bits() -> int:
    """format of the MLP backward's cos(u) dump (16 = fp16, default; 8 = CNG_G_DUMP_BITS=8, faster, 1/254 quantisation)"""
    from conditioned_nerf_gan_b200 import ops
    return ops.g_image_bytes() * 8 // (128 * 256)


def measure_c5(rays_m: int = 2):
    """BASELINE configs[4] (compositing + sample_pdf micro-benchmark): achieved HBM GB/s of the compositing, resampling and
    merge+composite kernels at 64 / 128 / 256 samples per ray on ``rays_m`` Mi rays, algorithmic bytes per ray from
    SURVEY.md 8(d), against the measured copy bandwidth.  Working sets (0.3 - 5 GB) exceed the 126 MB L2."""
    from conditioned_nerf_gan_b200 import ops
    from oracle import nerf_path as oracle
    peak = peaks()["hbm_gbs"]
    dev = "cuda"

    def timeit(fn, reps=5):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    rows = []
    n = rays_m << 20
    for S in (64, 128, 256):
        g = torch.Generator(device=dev).manual_seed(0)
        rs = torch.randn((n, S, 4), generator=g, device=dev)
        rs[..., :3].sigmoid_()
        t = torch.rand((n, S), generator=g, device=dev).mul_(1.7).add_(0.25).sort(dim=1).values
        w = torch.rand((n, S), generator=g, device=dev)
        u = torch.rand((n, S), generator=g, device=dev)
        rs2 = torch.randn((n, S, 4), generator=g, device=dev)
        t2 = torch.rand((n, S), generator=g, device=dev).mul_(1.7).add_(0.25)      # fine distances arrive UNSORTED (inverse-CDF of random u)
        rays = torch.nn.functional.normalize(torch.randn((1024, 3), generator=g, device=dev), dim=-1)
        cases = [("composite", lambda: ops.composite_fwd(rs, t, None, 0.0, "relu", True, False), n * (S * 20 + 16 + 4 * S)),
                 ("resample", lambda: ops.resample_from_coarse(t, w, u), n * 16 * S),
                 ("merge", lambda: ops.merge_composite(rs2, rs, t2, t, None, rays, n // 1024, 32, 32, 0.0, "relu", True, False), n * (2 * S * 20 + 16))]
        # same-box comparator: the reference's own eager op chain for the same step -- its torch-port restatement
        # (oracle/nerf_path.py: fancy_integration, the sample_pdf call site, cat + sort + gather) on cuda tensors, on a slice of
        # the rays (its temporaries are ~10x the inputs); a baseline leg, nothing of it is on the product path
        ne = min(n, 1 << 18)
        z4 = lambda a: a[:ne].unsqueeze(0)
        zero_noise = torch.zeros((1, ne, 2 * S, 1), device=dev)
        eager = {"composite": lambda: oracle.composite(z4(rs), z4(t).unsqueeze(-1), zero_noise[:, :, :S], 0.0, "relu", True, False),
                 "resample": lambda: oracle.coarse_to_fine_t(z4(w), z4(t), u[:ne], S),
                 "merge": lambda: oracle.composite(*oracle.merge_by_depth(z4(rs2), z4(rs), z4(t2).unsqueeze(-1), z4(t).unsqueeze(-1))[:2],
                                                   zero_noise, 0.0, "relu", True, False)}
        for name, fn, by in cases:
            ms = timeit(fn)
            ems = timeit(eager[name], reps=3) * (n / ne)
            rows.append({"kernel": name, "samples_per_ray": S if name != "merge" else f"{S}+{S}", "rays": n, "ms": ms, "gbs": by / ms / 1e6,
                         "frac": by / ms / 1e6 / peak, "gpu_eager_ms": ems, "speedup_vs_gpu_eager": ems / ms})
        del rs, t, w, u, rs2, t2, zero_noise
        torch.cuda.empty_cache()
    return {"peak_gbs": peak, "unit": "GB/s", "rays": n, "rows": rows,
            "bytes_per_ray": "composite S*20+16+4S (weights emitted), resample 16*S, merge 2S*20+16 (SURVEY.md 8d)",
            "gpu_eager": "oracle/nerf_path.py (torch port of fancy_integration / the sample_pdf call site / cat+sort+gather+fancy_integration) "
                         "run eagerly on cuda on 262144 of the rays, scaled to the full ray count"}


def measure_gpu_eager(args, steps=5, warmup=3):
    """The same-box GPU comparator BASELINE.md 5.3 asks for: the reference's eager PyTorch path on THIS B200 -- the torch
    port of it (oracle/nerf_path.py, every statement citing the reference line it restates; the reference itself cannot travel
    to the GPU box) on CUDA tensors, same workload, weights and draws -- in fp32 and under the trainer's autocast (fp16),
    CUDA-event timed.  A baseline leg: nothing of it is on the product path."""
    from oracle import nerf_path as oracle
    B, img, S, V = WORKLOAD["batch"], WORKLOAD["img_size"], WORKLOAD["num_steps"], WORKLOAD["volume"]
    dev = torch.device("cuda", torch.cuda.current_device())
    state = {k: v.to(dev) for k, v in oracle.init_generator_state(args.siren, seed=0).items()}
    vol, glob, cam = (t.to(dev) for t in synthetic_inputs(B, V, 0))
    meta = render_meta(img, S)
    g = torch.Generator().manual_seed(1)
    draws = {k: v.to(dev) for k, v in oracle.draw_randoms(B, img, S, True, g).items()}
    out = {"kind": "port", "what": "oracle/nerf_path.py (torch restatement of generators/generators.py:33-187) run eagerly on cuda, batch 8, same workload",
           "unit": UNIT, "steps": steps, "warmup": warmup}
    for name, amp in (("fp32", False), ("amp_fp16", True)):
        def fn():
            with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                return oracle.render(state, args.siren, (vol, glob), cam, draws, taps=False, **meta)
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
        out[name] = {"value": B * img * img / (ms * 1e-3), "ms_per_step": ms}
        torch.cuda.empty_cache()
    return out


def run_train_generator(args):
    """Generator-only train step (BASELINE config 3 minus the out-of-scope encoder / discriminator)."""
    import torch.distributed as dist
    rank, world, local, dev, barrier, max_over_ranks = _setup_ranks(args)
    GLOBAL_B, img, S, V = args.global_batch, 128, 48, 64
    if GLOBAL_B % world:
        raise SystemExit(f"global batch {GLOBAL_B} must divide by the number of GPUs")
    b = GLOBAL_B // world
    meta = render_meta(img, S)
    meta["nerf_noise"] = 0.5
    gen = random_init_generator(args.siren)
    gen = gen.to(dev)
    gen.set_device(dev)
    gen.siren.precision = args.precision
    model = gen
    if world > 1:
        model = torch.nn.parallel.DistributedDataParallel(gen, device_ids=[local], find_unused_parameters=True)   # utils.py:322-326
    opt = torch.optim.Adam(gen.parameters(), lr=5e-5, betas=(0.0, 0.9))
    vol_h, glob_h, cam_h = (t.pin_memory() for t in synthetic_inputs(b, V, seed=rank))
    target_h = (torch.rand((b, 3, img, img), generator=torch.Generator().manual_seed(100 + rank)) * 2 - 1).pin_memory()
    losses = []

    def step():
        vol = vol_h.to(dev, non_blocking=True).requires_grad_(True)      # stands for the encoder output: receives a gradient
        glob = glob_h.to(dev, non_blocking=True).requires_grad_(True)
        cam, target = cam_h.to(dev, non_blocking=True), target_h.to(dev, non_blocking=True)
        opt.zero_grad(set_to_none=True)
        pixels, depth = model((vol, glob), cam, **meta)
        loss = torch.nn.functional.mse_loss(pixels, target) + 0.1 * depth.mean()         # photo + depth terms (utils.py:673-706)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(gen.parameters(), 1.0)                               # utils.py:726-729
        opt.step()
        losses.append(loss.detach())

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = _timed_steps(step, args.steps, args.warmup, barrier, max_over_ranks)
    clocks = sampler.stop() if sampler else None
    final_loss = float(losses[-1].item())        # device->host read of the step's result
    if rank == 0:
        L = SIREN_LAYERS[args.siren]
        pts = GLOBAL_B * img * img * 2 * S
        line = {"metric": "train_images_per_sec_128x128_48+48spp_global_batch32_generator_only", "value": GLOBAL_B * args.steps / (ms * 1e-3),
                "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
                "config": {"workload": "generator_train_step_128x128_48+48spp_global_batch32 (BASELINE configs[2] without the out-of-scope U-Net encoder and discriminator)",
                           "siren_type": args.siren, "batch_per_gpu": b, "optimizer": "Adam", "grad_allreduce": "DDP/NCCL" if world > 1 else "none",
                           "backward": "hand-written kernels throughout: compositing / scatter, training-mode forward (recompute), tcgen05 dgrad chain, tcgen05 split-K weight gradient"},
                "e2e": {"value": GLOBAL_B * args.steps / (ms * 1e-3), "unit": "images/s",
                        "h2d_bytes_per_step": int((vol_h.numel() + glob_h.numel() + cam_h.numel() + target_h.numel()) * 4), "d2h_bytes_per_step": 4},
                "gpu_launches": launches, "clocks": clocks, "final_loss": final_loss,
                "mlp_flops_per_step_fwd": mlp_flops_per_point(L) * pts}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_video(args):
    """BASELINE config 4: 64 poses of one object, 256x256, 48+48 samples, poses sharded over the ranks."""
    import torch.distributed as dist
    from conditioned_nerf_gan_b200 import parallel
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import create_cam2world_matrix
    rank, world, local, dev, barrier, max_over_ranks = _setup_ranks(args)
    P, img, S, V = 64, 256, 48, 64
    gen = random_init_generator(args.siren)
    gen = gen.to(dev).eval()
    gen.set_device(dev)
    gen.siren.precision = args.precision
    vol, glob, _ = synthetic_inputs(1, V, seed=0)
    z = (vol.to(dev), glob.to(dev))
    parallel.broadcast_z(z, src=0)
    # camera spiral around the object (inference.py:442-477 style), fov sweep 60 -> 30 (inference.py:459)
    k = torch.arange(P, dtype=torch.float32) / P
    theta, phi, r = 2 * np.pi * k, 0.35 * np.pi + 0.15 * np.pi * torch.sin(2 * np.pi * k), 1.2
    origin = torch.stack([r * torch.sin(phi) * torch.cos(theta), r * torch.cos(phi), r * torch.sin(phi) * torch.sin(theta)], -1)
    poses = create_cam2world_matrix(origin, "y", "cpu").to(dev)
    fov = [60.0 - 30.0 * (i // 8) / (P // 8 - 1) for i in range(P)]        # piecewise constant so that chunks of 8 share a fov
    meta = {kk: v for kk, v in render_meta(img, S).items() if kk != "fov"}
    frames_h = torch.empty((P, 3, img, img), dtype=torch.float32).pin_memory() if rank == 0 else None

    def step():
        pixels, depth = parallel.render_poses_sharded(gen, z, poses, fov=fov, max_batch_size=4, **meta)
        if rank == 0:
            frames_h.copy_(pixels, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    ms, launches = _timed_steps(step, args.steps, args.warmup, barrier, max_over_ranks)
    clocks = sampler.stop() if sampler else None
    if rank == 0:
        rays = P * img * img * args.steps
        line = {"metric": "video_rays_per_sec_256x256_48+48spp_64poses", "value": rays / (ms * 1e-3), "unit": "rays/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
                "config": {"workload": "staged_forward video render 256x256, 48+48 samples, 64 poses, one 64^3x32 object (BASELINE configs[3])",
                           "siren_type": args.siren, "frames_per_s": P * args.steps / (ms * 1e-3), "poses_per_gpu": P // world,
                           "parallelism": f"poses sharded over {world} GPU(s); one all_gather of frames per step"},
                "e2e": {"value": rays / (ms * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": int(P * 3 * img * img * 4)},
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--siren", default="TALLSIREN_FG", choices=sorted(SIREN_LAYERS))
    ap.add_argument("--precision", default=None, choices=["bf16", "fp16", "fp32"],
                    help="default: bf16 operands; fp16 for the classes offered with fp16 operands only (SHORTSIREN_FG)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-volume", default="fp16", choices=["fp16", "fp32"],
                    help="dtype of the feature volumes in host memory for the end-to-end leg (fp16 = the encoder's autocast output dtype)")
    ap.add_argument("--no-train", action="store_true", help="skip the config-3 train-step leg of the default line")
    ap.add_argument("--train-steps", type=int, default=4)
    ap.add_argument("--global-batch", type=int, default=32, help="train legs: global batch (BASELINE configs[2]: 32; other values are experiments)")
    ap.add_argument("--no-extras", action="store_true", help="skip the c5 micro-benchmark table and the GPU-eager comparator")
    ap.add_argument("--workload", default="render", choices=["render", "train", "train_generator", "video"])
    args = ap.parse_args()
    if args.precision is None:
        args.precision = "fp16" if args.siren.startswith("SHORT") else "bf16"
    if args.warmup < 3 and args.impl == "b200":
        print(f"note: --warmup {args.warmup} < 3 (timing rules ask for >= 3)", file=sys.stderr)
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train_generator":
        run_train_generator(args)
    elif args.workload == "train":
        run_train(args)
    elif args.workload == "video":
        run_video(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
