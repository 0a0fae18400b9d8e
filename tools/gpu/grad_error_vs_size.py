"""Gradient error of the CUDA backward against torch autograd through the oracle as a function of the scene size, in both formats
of the cos(u) dump (fp16 / 8-bit codes): TALLSIREN_FG with density, batch 1, 12+12 samples, 16^3 x 32 volume.
    python tools/gpu/grad_error_vs_size.py [sizes ...]"""
import ctypes, os, sys
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
from conditioned_nerf_gan_b200 import _lib
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
from oracle import nerf_path as oracle

lib = _lib.load()
lib.cng_internal_set_g_dump_bits.argtypes = [ctypes.c_int]
lib.cng_internal_set_g_dump_bits.restype = None
sizes = [int(a) for a in sys.argv[1:]] or [8, 16, 32, 64]
S, V, B = 12, 16, 1
state = oracle.init_generator_state("TALLSIREN_FG", 256, 32, 256, seed=1)     # random init, as the gradient fixtures of tests/
dev = torch.device("cuda")
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-300))
for img in sizes:
    g = torch.Generator().manual_seed(2)
    volume = torch.randn((B, 32, V, V, V), generator=g) * 0.3
    glob = torch.randn((B, 256), generator=g) * 0.05 + 0.19
    cam = oracle.look_at_cam2world(oracle.random_camera_origins(B, 0.7, 1.5, "y", np.random.RandomState(3)), "y")
    draws = oracle.draw_randoms(B, img, S, True, g)
    meta = dict(img_size=img, fov=49.134342641202636, ray_start=0.25, ray_end=1.95, num_steps=S, hierarchical_sample=True,
                clamp_mode="relu", nerf_noise=0.0, white_back=True)
    d_pix, d_dep = torch.randn((B, 3, img, img), generator=g), torch.randn((B, img, img), generator=g)
    st = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    vol_r, glob_r = volume.clone().requires_grad_(True), glob.clone().requires_grad_(True)
    out = oracle.render_with_grad(st, "TALLSIREN_FG", (vol_r, glob_r), cam, draws, **meta)
    ((out["pixels"] * d_pix).sum() + (out["depth"] * d_dep).sum()).backward()
    ref = {k: v.grad for k, v in st.items()}
    ref["volume"], ref["global"] = vol_r.grad, glob_r.grad
    res = {}
    for bits in (16, 8):
        lib.cng_internal_set_g_dump_bits(bits)
        gen = ImplicitGenerator3d("TALLSIREN_FG", 256, 32, 4, 256)
        gen.load_state_dict(state, strict=True)
        gen = gen.to(dev)
        gen.set_device(dev)
        gen.siren.precision = "fp32"            # exact forward (as tests/test_gpu_backward.py): what is measured is the backward
        vol = volume.to(dev).requires_grad_(True)
        gl = glob.to(dev).requires_grad_(True)
        pixels, depth = gen((vol, gl), cam.to(dev), draws={k: v.to(dev) for k, v in draws.items()}, **meta)
        ((pixels * d_pix.to(dev)).sum() + (depth * d_dep.to(dev)).sum()).backward()
        torch.cuda.synchronize()
        got = {"siren." + k: p.grad.cpu() for k, p in gen.siren.named_parameters()}
        got["volume"], got["global"] = vol.grad.cpu(), gl.grad.cpu()
        res[bits] = got
    lib.cng_internal_set_g_dump_bits(0)
    worst = {b: max(rel(res[b][k], ref[k]) for k in ref if k in res[b]) for b in (16, 8)}
    between = max(rel(res[8][k], res[16][k]) for k in res[16])
    w0 = {b: rel(res[b]["siren.network.0.layer.weight"], ref["siren.network.0.layer.weight"]) for b in (16, 8)}
    groups = {"hidden weights": [k for k in ref if "network" in k and k.endswith("weight") and ".0." not in k], "layer-0 weight": ["siren.network.0.layer.weight"],
              "biases": [k for k in ref if "network" in k and k.endswith("bias")], "head": [k for k in ref if "final_layer" in k],
              "mapping network": [k for k in ref if "mapping" in k], "volume": ["volume"], "global feature": ["global"]}
    for name, keys in groups.items():
        print(f"      {name:16s} vs oracle: fp16 {max(rel(res[16][k], ref[k]) for k in keys):.2e}  8-bit {max(rel(res[8][k], ref[k]) for k in keys):.2e}   "
              f"8-bit vs fp16 {max(rel(res[8][k], res[16][k]) for k in keys):.2e}")
    print(f"{img:3d} x {img:<3d} ({img * img * 2 * S:7d} points): worst rel-L2 vs oracle  fp16 cos {worst[16]:.3e}   8-bit cos {worst[8]:.3e}   "
          f"8-bit vs fp16 {between:.3e}   layer-0 weight {w0[16]:.2e} / {w0[8]:.2e}")
