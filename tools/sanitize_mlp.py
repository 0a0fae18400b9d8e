#!/usr/bin/env python
"""Tiny K2 launches: every tcgen05 kernel version, plain and residual networks, inference and training mode, each against
the exact fp32 kernel.  Written for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_mlp.py`; that tool is
closed on this GPU pool, so it runs as a plain all-versions smoke (profiles/r1h_k2_all_versions_smoke.txt)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conditioned_nerf_gan_b200 import _lib, ops
from oracle import nerf_path as oracle

lib = _lib.load()
lib.cng_internal_set_tc_version.argtypes = [ctypes.c_int]
lib.cng_internal_set_tc_version.restype = None
dev = "cuda"
for siren, B, N in (("TALLSIREN_FG", 2, 128 * 5 + 7), ("SingleSIREN_dg", 1, 300), ("TALLSIREN_dRes", 2, 128 * 3 + 1)):
    spec = oracle.SIREN_SPECS[siren]
    st = oracle.init_generator_state(siren, seed=0)
    ws = [st[f"siren.{k}.weight"].to(dev) for k in oracle.layer_keys(siren)]
    bs = [st[f"siren.{k}.bias"].to(dev) for k in oracle.layer_keys(siren)]
    L = len(ws)
    g = torch.Generator().manual_seed(1)
    if spec.get("film", True):
        glob = torch.randn((B, 256), generator=g) * 0.05 + 0.19
        freq, phase = (t.to(dev) for t in oracle.film_parameters(glob, st["siren.mapping_network.weight"], st["siren.mapping_network.bias"]))
    else:
        freq, phase = torch.ones((B, L * 256), device=dev), torch.zeros((B, L * 256), device=dev)
    feat = (torch.randn((B, N, 32), generator=g) * 0.3).to(dev)
    fw, fb = st["siren.final_layer.weight"].to(dev), st["siren.final_layer.bias"].to(dev)
    ref = ops.film_siren_fwd(feat, ws, bs, freq, phase, fw, fb, spec["sigmoid_rgb"], "fp32", spec.get("res_save", 0), spec.get("res_add", 0))
    for v in (1, 2, 3):
        lib.cng_internal_set_tc_version(v)
        out = ops.film_siren_fwd(feat, ws, bs, freq, phase, fw, fb, spec["sigmoid_rgb"], "bf16", spec.get("res_save", 0), spec.get("res_add", 0))
        torch.cuda.synchronize()
        print(f"{siren} v{v}: max-abs vs fp32 kernel {(out - ref).abs().max().item():.3e}", flush=True)
    lib.cng_internal_set_tc_version(0)
    if siren != "SingleSIREN_dg":
        _, xs, gs = ops.film_siren_fwd_train(feat, ws, bs, freq, phase, fw, fb, spec["sigmoid_rgb"], spec.get("res_save", 0), spec.get("res_add", 0))
        torch.cuda.synchronize()
        print(f"{siren} train-mode dumps: x {tuple(xs.shape)} finite={bool(torch.isfinite(xs.float()).all())}", flush=True)
