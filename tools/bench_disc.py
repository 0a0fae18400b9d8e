#!/usr/bin/env python
"""Discriminator update with the R1 penalty (double backward through the convolutions), batch 8 at 128x128: cuDNN algorithm
selection (cudnn.benchmark) x memory format x autocast dtype.   python tools/bench_disc.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator

dev = torch.device("cuda")
b = 8
real = torch.rand((b, 3, 128, 128), device=dev) * 2 - 1
fake = torch.rand((b, 3, 128, 128), device=dev) * 2 - 1

def run(cl, bench, dtype):
    torch.backends.cudnn.benchmark = bench
    disc = ProgressiveDiscriminator().to(dev)
    if cl:
        disc = disc.to(memory_format=torch.channels_last)
    opt = torch.optim.Adam(disc.parameters(), lr=1e-4)
    def step():
        r = real.clone()
        f = fake
        if cl:
            r, f = r.contiguous(memory_format=torch.channels_last), f.contiguous(memory_format=torch.channels_last)
        r.requires_grad_(True)
        with torch.autocast("cuda", dtype=dtype):
            rp = disc(r, 1.0)
        g = torch.autograd.grad(rp.sum() * 1024.0, r, create_graph=True)[0] / 1024.0
        with torch.autocast("cuda", dtype=dtype):
            pen = 5.0 * (g.reshape(b, -1).norm(2, dim=1) ** 2).mean()
            loss = F.softplus(disc(f, 1.0)).mean() + F.softplus(-rp).mean() + pen
        opt.zero_grad()
        (loss * 1024.0).backward()
        opt.step()
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5):
        step()
    e.record(); torch.cuda.synchronize()
    print(f"channels_last={cl} cudnn.benchmark={bench} {dtype}: D update {s.elapsed_time(e) / 5:.1f} ms (batch {b})", flush=True)

for dtype in (torch.float16, torch.bfloat16):
    for cl in (False, True):
        for bench in (False, True):
            run(cl, bench, dtype)
