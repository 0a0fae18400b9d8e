#!/usr/bin/env python
"""Eager vs CUDA-graph forward at the launch-bound BASELINE config 1 (64x64, 12+12, batch 1, 32^3 volume)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
from conditioned_nerf_gan_b200.graphs import GraphedRender
from oracle import nerf_path as oracle

dev = torch.device("cuda")
for name, B, img, S, V in (("c1: B1 64x64 12+12 V32", 1, 64, 12, 32), ("B1 128x128 24+24 V64", 1, 128, 24, 64)):
    gen = ImplicitGenerator3d("TALLSIREN_FG", 256, 32, 4, 256)
    gen.load_state_dict(oracle.init_generator_state("TALLSIREN_FG", seed=0), strict=True)
    gen = gen.to(dev).eval()
    vol, glob, cam = (t.to(dev) for t in bench.synthetic_inputs(B, V, 0))
    meta = bench.render_meta(img, S)
    render = GraphedRender(gen, (vol, glob), cam, **meta)

    def timeit(fn, reps=200):
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    with torch.no_grad():
        eager = timeit(lambda: gen((vol, glob), cam, **meta))
    graphed = timeit(lambda: render((vol, glob), cam))
    print(f"{name}: eager {eager:.3f} ms ({B * img * img / eager / 1e3:.2f} M rays/s), CUDA graph {graphed:.3f} ms ({B * img * img / graphed / 1e3:.2f} M rays/s)")
