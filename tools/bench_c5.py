#!/usr/bin/env python
"""BASELINE configs[4] on its own: compositing / resampling / merge+composite at 64 / 128 / 256 samples per ray (bench.measure_c5).
    python tools/bench_c5.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench

r = bench.measure_c5()
print(f'# {r["rays"]} rays, peak {r["peak_gbs"]} GB/s; bytes per ray: {r["bytes_per_ray"]}')
for row in r["rows"]:
    print(f'{row["kernel"]:10s} {str(row["samples_per_ray"]):8s} {row["ms"]:8.3f} ms {row["gbs"]:8.0f} GB/s {row["frac"] * 100:5.1f} %   '
          f'GPU-eager reference path {row["gpu_eager_ms"]:9.2f} ms ({row["speedup_vs_gpu_eager"]:.0f}x)')
