#!/usr/bin/env python
"""Kernel-time breakdown of one generator train step (forward with grad + backward) via torch.profiler."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
from oracle import nerf_path as oracle

b = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda")
gen = ImplicitGenerator3d("TALLSIREN_FG", 256, 32, 4, 256)
gen.load_state_dict(oracle.init_generator_state("TALLSIREN_FG", seed=0), strict=True)
gen = gen.to(dev)
meta = bench.render_meta(128, 48)
vol, glob, cam = (t.to(dev) for t in bench.synthetic_inputs(b, 64, 0))
target = torch.rand((b, 3, 128, 128), device=dev) * 2 - 1

def step():
    v = vol.clone().requires_grad_(True)
    g = glob.clone().requires_grad_(True)
    gen.zero_grad(set_to_none=True)
    pixels, depth = gen((v, g), cam, **meta)
    (torch.nn.functional.mse_loss(pixels, target) + 0.1 * depth.mean()).backward()

step(); torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record(); step(); e.record(); torch.cuda.synchronize()
print(f"batch {b}: {s.elapsed_time(e):.1f} ms per step = {s.elapsed_time(e) / b:.2f} ms per image")
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=28, max_name_column_width=70))
