#!/usr/bin/env python
"""Micro-benchmark of cng_film_siren_fwd alone at the c2 shape (B=8, N=128*128*24 points, L layers).
    python tools/bench_mlp.py [SIREN_TYPE] [reps]        (CNG_TC_POLY selects the sine split of the tcgen05 kernel)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conditioned_nerf_gan_b200 import ops
from oracle import nerf_path as oracle

siren = sys.argv[1] if len(sys.argv) > 1 else "TALLSIREN_FG"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
prec = sys.argv[3] if len(sys.argv) > 3 else "bf16"
B, N = 8, 128 * 128 * 24
spec = oracle.SIREN_SPECS[siren]
L = spec["layers"]
st = oracle.init_generator_state(siren, seed=0)
dev = "cuda"
ws = [st[f"siren.network.{i}.layer.weight"].to(dev) for i in range(L)]
bs = [st[f"siren.network.{i}.layer.bias"].to(dev) for i in range(L)]
g = torch.Generator().manual_seed(1)
glob = torch.randn((B, 256), generator=g) * 0.05 + 0.19
freq, phase = oracle.film_parameters(glob, st["siren.mapping_network.weight"], st["siren.mapping_network.bias"])
freq, phase = freq.to(dev), phase.to(dev)
feat = (torch.randn((B, N, 32), generator=g) * 0.3).to(dev)
fw, fb = st["siren.final_layer.weight"].to(dev), st["siren.final_layer.bias"].to(dev)
run = lambda prec: ops.film_siren_fwd(feat, ws, bs, freq, phase, fw, fb, spec["sigmoid_rgb"], prec)
out = run(prec)
ref = run("fp32")
torch.cuda.synchronize()
err = (out - ref).abs().max().item()
for _ in range(3):
    run(prec)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(reps):
    run(prec)
e.record()
torch.cuda.synchronize()
ms = s.elapsed_time(e) / reps
flops = 2 * (32 * 256 + (L - 1) * 256 * 256 + 256 * 4) * B * N
chk = int(out.view(torch.int32).long().sum().item())
print(f"{siren} {prec} v={os.environ.get('CNG_TC_V', 'default')} bits-checksum {chk} poly={os.environ.get('CNG_TC_POLY', 'default')}: {ms:.3f} ms/launch, {flops / ms / 1e9:.1f} TFLOP/s algorithmic, "
      f"max-abs vs fp32 kernel {err:.3e}")
