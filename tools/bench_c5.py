#!/usr/bin/env python
"""BASELINE config 5: compositing + sample_pdf micro-benchmark sweep, achieved HBM GB/s against the measured
copy bandwidth (MEASURED_PEAKS.json).  Algorithmic bytes per ray:
  composite (weights emitted)   S*(16+4) read + 16 + 4*S written
  merge_composite               2S*(16+4) read + 16 written (pixels + depth)
  sample_pdf                    4*((M+1) + M + K) read + 4*K written
    python tools/bench_c5.py [--rays-max 16] [--json out.json]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conditioned_nerf_gan_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--rays-max", type=int, default=16, help="largest ray count in millions")
ap.add_argument("--json", default=None)
args = ap.parse_args()
peak = 6556.5
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk))["hbm_gbs"]
dev = "cuda"


def timeit(fn, reps=10):
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
for S in (64, 128, 256):
    for mrays in (1, 2, 4, 8, 16):
        if mrays > args.rays_max:
            continue
        n = mrays << 20
        if n * S * 28 > 60e9:           # keep the working set well inside HBM
            continue
        g = torch.Generator(device=dev).manual_seed(0)
        rs = torch.randn((n, S, 4), generator=g, device=dev)
        rs[..., :3].sigmoid_()
        t = torch.rand((n, S), generator=g, device=dev).mul_(1.7).add_(0.25).sort(dim=1).values
        ms = timeit(lambda: ops.composite_fwd(rs, t, None, 0.0, "relu", True, False))
        by = n * (S * 20 + 16 + 4 * S)
        rows.append(dict(kernel="composite_fwd", S=S, rays=n, ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / peak))
        # sample_pdf: M = S - 2 bins interior weights, K = S samples (the generator's call shape at num_steps = S)
        w = torch.rand((n, S), generator=g, device=dev)
        u = torch.rand((n, S), generator=g, device=dev)
        ms = timeit(lambda: ops.resample_from_coarse(t, w, u))
        by = n * (4 * (S + S + S) + 4 * S)
        rows.append(dict(kernel="resample_from_coarse", S=S, rays=n, ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / peak))
        del w, u
        if 2 * S <= 512 and n * S * 40 < 60e9:
            rs2 = torch.randn((n, S, 4), generator=g, device=dev)
            t2 = torch.rand((n, S), generator=g, device=dev).mul_(1.7).add_(0.25).sort(dim=1).values
            rays = torch.nn.functional.normalize(torch.randn((1024, 3), generator=g, device=dev), dim=-1)
            B, R = n // 1024, 1024
            ms = timeit(lambda: ops.merge_composite(rs2, rs, t2, t, None, rays, B, 32, 32, 0.0, "relu", True, False))
            by = n * (2 * S * 20 + 16)
            rows.append(dict(kernel="merge_composite", S=S, rays=n, ms=ms, gbs=by / ms / 1e6, frac=by / ms / 1e6 / peak))
            del rs2, t2
        del rs, t
        torch.cuda.empty_cache()
print(f"# HBM peak (measured copy) {peak} GB/s")
for r in rows:
    print(f"{r['kernel']:22s} S={r['S']:3d} rays={r['rays'] >> 20:2d}M  {r['ms']:8.3f} ms  {r['gbs']:7.1f} GB/s  {100 * r['frac']:5.1f}% of peak")
if args.json:
    json.dump(dict(peak_gbs=peak, rows=rows), open(args.json, "w"), indent=1)
