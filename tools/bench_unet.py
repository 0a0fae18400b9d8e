#!/usr/bin/env python
"""cuDNN algorithm / layout sweep for the 3D U-Net encoder (forward + backward, batch 8, 64^3): which memory format and
cudnn.benchmark setting the library runs fastest with.   python tools/bench_unet.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conditioned_nerf_gan_b200.generators import unet3d

dev = torch.device("cuda")
b = 8
x = torch.rand((b, 4, 64, 64, 64), device=dev)

def run(cl, bench, dtype):
    torch.backends.cudnn.benchmark = bench
    unet3d.FORCE_CONTIGUOUS = not cl
    enc = unet3d.UNet3D(in_channels=4, out_channels=32, f_maps=32, num_levels=4, is_segmentation=False, final_sigmoid=False, return_global=True).to(dev)
    def step():
        with torch.autocast("cuda", dtype=dtype):
            fv, g = enc(x)
            loss = fv.float().square().mean() + g.float().mean()
        loss.backward()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        step()
    e.record(); torch.cuda.synchronize()
    with torch.no_grad(), torch.autocast("cuda", dtype=dtype):
        for _ in range(2): enc(x)
        torch.cuda.synchronize(); s2, e2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s2.record()
        for _ in range(5): enc(x)
        e2.record(); torch.cuda.synchronize()
    print(f"channels_last_3d={cl} cudnn.benchmark={bench} {dtype}: fwd+bwd {s.elapsed_time(e) / 5:.1f} ms, fwd {s2.elapsed_time(e2) / 5:.1f} ms (batch {b})", flush=True)

for dtype in (torch.float16, torch.bfloat16):
    for cl in (True, False):
        for bench in (False, True):
            run(cl, bench, dtype)
