#!/usr/bin/env python
"""Forward timing of the generator across BASELINE configs 1-2 and the SIREN variants (inputs resident, CUDA events).
    python tools/bench_configs.py > gpurun_out/configs.log"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
from oracle import nerf_path as oracle

dev = torch.device("cuda")
CONFIGS = [("c1: B1 64x64 12+12 V32", 1, 64, 12, 32), ("c2: B8 128x128 24+24 V64", 8, 128, 24, 64),
           ("B1 128x128 48+48 V64", 1, 128, 48, 64), ("B4 256x256 48+48 V64", 4, 256, 48, 64)]
print(f"{'config':28s} {'siren':16s} {'prec':5s} {'ms':>8s} {'M rays/s':>9s} {'MLP TFLOP/s (algorithmic, whole step)':>10s}")
for name, B, img, S, V in CONFIGS:
    for siren in ("TALLSIREN_FG", "SHORTSIREN_FG", "DOUBLESIREN_FG", "SingleSIREN_dg", "SHORTSIREN_F", "TALLSIREN_dRes", "TALLSIREN_dResLong",
                  "SHORTSIREN_FRes"):
        spec = oracle.SIREN_SPECS[siren]
        gen = ImplicitGenerator3d(siren, 32 if siren.startswith("TALLSIREN_dRes") else 256, 32, 4, 256)
        gen.load_state_dict(oracle.init_generator_state(siren, seed=0), strict=True)
        gen = gen.to(dev).eval()
        vol, glob, cam = (t.to(dev) for t in bench.synthetic_inputs(B, V, 0))
        z = (vol, glob) if spec.get("film", True) else vol
        meta = bench.render_meta(img, S)
        for prec in (["bf16", "fp16", "fp32"] if (siren == "TALLSIREN_FG" and B * img * img <= 8 * 128 * 128) else [gen.siren.precision]):
            gen.siren.precision = prec
            with torch.no_grad():
                for _ in range(3):
                    gen(z, cam, **meta)
                torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = 10 if prec != "fp32" else 3
                s.record()
                for _ in range(reps):
                    gen(z, cam, **meta)
                e.record()
                torch.cuda.synchronize()
            ms = s.elapsed_time(e) / reps
            rays = B * img * img
            flops = bench.mlp_flops_per_point(spec["layers"]) * rays * 2 * S
            print(f"{name:28s} {siren:16s} {prec:5s} {ms:8.3f} {rays / ms / 1e3:9.2f} {flops / ms / 1e9:10.1f}")
