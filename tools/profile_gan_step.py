#!/usr/bin/env python
"""Kernel-time breakdown of one full GAN train step (U-Net encoder + generator + discriminator) via torch.profiler, plus
CUDA-event timings of the phases.   python tools/profile_gan_step.py [batch]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
from conditioned_nerf_gan_b200.generators.unet3d import UNet3D
from conditioned_nerf_gan_b200.generators.volumetric_rendering import create_cam2world_matrix, sample_camera_positions
from conditioned_nerf_gan_b200.training import GanTrainStep

b = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
torch.manual_seed(0); np.random.seed(0)
gen = ImplicitGenerator3d("TALLSIREN_FG", 256, 32, 4, 256).to(dev)
enc = UNet3D(in_channels=4, out_channels=32, f_maps=32, num_levels=4, is_segmentation=False, final_sigmoid=False, return_global=True).to(dev)
disc = ProgressiveDiscriminator().to(dev)
md = dict(bench.render_meta(128, 48), nerf_noise=1.0, batch_split=1, r1_lambda=10, grad_clip=1, betas=(0.0, 0.9), weight_decay=0, gen_lr=1e-5,
          disc_lr=1e-4, enc_lr=2e-5, photo_loss=True, depth_loss=False, enable_discriminator=True, random_gen_img=True, cam_r_start=0.7, cam_r_end=1.5)
tr = GanTrainStep(gen, enc, disc, md, dev, amp=True)
occ = (torch.rand((b, 1, 64, 64, 64), device=dev) < 0.05).float()
sample = {"img": torch.rand((b, 3, 128, 128), device=dev) * 2 - 1, "voxel": torch.cat([occ, torch.rand((b, 3, 64, 64, 64), device=dev) * occ], 1),
          "cam2world": create_cam2world_matrix(sample_camera_positions(dev, "y", 0.7, 1.5, b), "y", dev)}

def timed(fn):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record(); r = fn(); e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e), r

for _ in range(2):
    tr.step(sample)
print(f"batch {b}: full step {timed(lambda: tr.step(sample))[0]:.1f} ms; D step {timed(lambda: tr.train_discriminator(sample))[0]:.1f} ms; "
      f"G/E step {timed(lambda: tr.train_generator(sample))[0]:.1f} ms")
with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
    t_enc, z = timed(lambda: enc(sample["voxel"]))
    t_gen, out = timed(lambda: gen(z, sample["cam2world"], **tr.metadata))
    t_disc, _ = timed(lambda: disc(out[0], 1.0))
print(f"no-grad forwards: encoder {t_enc:.1f} ms, generator {t_gen:.1f} ms, discriminator {t_disc:.1f} ms")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    tr.step(sample); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=80))
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=25, max_name_column_width=80))
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); tr.step(sample); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host time to issue one step {1e3 * (t1 - t0):.1f} ms; until the device is done {1e3 * (t2 - t0):.1f} ms")
