#!/usr/bin/env python
"""Per-kernel timing of the tcgen05 MLP backward (a14) on one chunk of points: training-mode forward (recompute with dumps),
dgrad chain, split-K weight gradient, and the one-call cng_film_siren_bwd.  Algorithmic bytes per point and layer: recompute
512 + G B written (x + g), dgrad G + 512 B (g read, dz written), wgrad 1 KB (dz + x read), G = 512 (fp16 cos, default) or 256 (CNG_G_DUMP_BITS=8); FLOPs per point and hidden layer 2*256^2 each.
    python tools/bench_bwd.py [--points 1048576] [--siren TALLSIREN_FG]"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from conditioned_nerf_gan_b200 import ops
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d

ap = argparse.ArgumentParser()
ap.add_argument("--points", type=int, default=1 << 20)
ap.add_argument("--siren", default="TALLSIREN_FG")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda")
torch.manual_seed(0)
gen = ImplicitGenerator3d(args.siren, 256, 32, 4, 256).to(dev)
net = gen.siren
ws, bs = [w.detach() for w in net.layer_parameters()[0]], [b.detach() for b in net.layer_parameters()[1]]
L, P = len(ws), args.points
feat = torch.randn((1, P, 32), device=dev) * 0.3
glob = torch.randn((1, 256), device=dev) * 0.05 + 0.19
freq, phase = net.film_parameters(glob, 1, dev)
fw, fb = net.final_layer.weight.detach(), net.final_layer.bias.detach()
d_out = torch.randn((P, 4), device=dev)


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(args.reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / args.reps


hidden_flops = 2 * 256 * 256 * (L - 1) * P
G = ops.g_image_bytes() // 128          # bytes of cos(u) per point and layer
rows = []
ms = timeit(lambda: ops.film_siren_fwd(feat, ws, bs, freq, phase, fw, fb, net.sigmoid_rgb, "fp16"))
rows.append(("inference forward (fp16 operands)", ms, None, hidden_flops))
out, xs, gs, fd = ops.film_siren_fwd_train(feat, ws, bs, freq, phase, fw, fb, net.sigmoid_rgb, "fp16")
ms = timeit(lambda: ops.film_siren_fwd_train(feat, ws, bs, freq, phase, fw, fb, net.sigmoid_rgb, "fp16"))
rows.append(("training forward (recompute + dumps)", ms, P * L * (512 + G) + P * 128 * 2, hidden_flops))
wt = ops.film_siren_wt_images(ws, fw)
d_fb = torch.zeros(4, device=dev)
d_feat, dz = ops.film_siren_dgrad(d_out, out[0], net.sigmoid_rgb, L, wt, gs, d_fb)
ms = timeit(lambda: ops.film_siren_dgrad(d_out, out[0], net.sigmoid_rgb, L, wt, gs, d_fb))
rows.append(("dgrad chain", ms, P * L * (512 + G) + P * 160, hidden_flops))
dW = [torch.zeros_like(w) for w in ws]
colsum = torch.zeros((L, 256), device=dev)
ms = timeit(lambda: ops.film_siren_wgrad(dz, xs, fd, P, L, True, dW, colsum))
rows.append(("weight gradient (split-K)", ms, P * L * 1024, hidden_flops))
del xs, gs, fd, dz
torch.cuda.empty_cache()
d_fw, d_feat2 = torch.zeros_like(fw), torch.empty((P, 32), device=dev)
ms = timeit(lambda: ops.film_siren_bwd(feat[0], d_out, ws, bs, freq[0].contiguous(), phase[0].contiguous(), fw, fb, net.sigmoid_rgb, d_feat2, dW, colsum, d_fw, d_fb))
rows.append(("cng_film_siren_bwd (all of the above but the inference forward)", ms, P * L * (2 * (512 + G) + 1024), 3 * hidden_flops))
print(f"# {args.siren} L={L}, {P} points per chunk, cos dump {8 * G // 256} bits")
for name, ms, by, fl in rows:
    gbs = f"{by / ms / 1e6:7.0f} GB/s" if by else "            "
    print(f"{name:66s} {ms:8.3f} ms  {gbs}  {fl / ms / 1e9:7.1f} TFLOP/s (hidden layers)")
