"""Small, odd-sized calls of the compositing / merge / resampling kernels and the MLP backward, to run under compute-sanitizer:
    compute-sanitizer --tool memcheck python tools/gpu/sanitize_c5.py"""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from conditioned_nerf_gan_b200 import ops
from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
for S in (1, 2, 3, 12, 16, 17, 24, 31, 32, 33, 48, 63, 64, 65, 100, 127, 128, 129, 200, 255, 256):
    img = 3
    R = img * img
    fine = torch.randn((2, R, S, 4), generator=g, device=dev)
    coarse = torch.randn((2, R, S, 4), generator=g, device=dev)
    t_c = (torch.rand((2, R, S, 1), generator=g, device=dev) * 1.7 + 0.25).sort(dim=2).values
    t_f = torch.rand((2, R, S, 1), generator=g, device=dev) * 1.7 + 0.25
    if S > 2:
        t_c[0, 1] = t_c[0, 1].flip(0)          # the generic path
        t_f[1, 0, 0] = float("nan")            # a NaN distance must stay in bounds
    noise = torch.randn((2, R, 2 * S, 1), generator=g, device=dev)
    rays, _ = camera_tables((img, img), S, 49.0, 0.25, 1.95, dev)
    ops.merge_composite(fine, coarse, t_f, t_c, noise, rays, 2, img, img, 0.5, "softplus", False, True, taps=True)
    ops.merge_composite(None, coarse, None, t_c, None, rays, 2, img, img, 0.0, "relu", True, False)
    ops.merge_sort(t_f.squeeze(-1).reshape(2 * R, S), t_c.squeeze(-1).reshape(2 * R, S), want_sorted=True) if S > 1 else None
    ops.composite_fwd(coarse, t_c, noise[:, :, :S], 0.3, "relu", True, True)
    if S >= 3:
        ops.resample_from_coarse(t_c.squeeze(-1).reshape(2 * R, S), torch.rand((2 * R, S), generator=g, device=dev),
                                 torch.rand((2 * R, S), generator=g, device=dev))
for S in (300, 512, 1000):
    x = torch.randn((1, 5, S, 4), generator=g, device=dev)
    t = torch.rand((1, 5, S, 1), generator=g, device=dev).sort(dim=2).values
    ops.composite_fwd(x, t, None, 0.0, "relu", True, False)
for n, M, K in ((7, 1, 5), (5, 30, 32), (5, 31, 33), (5, 62, 64), (5, 63, 64), (3, 126, 128), (3, 127, 128), (3, 254, 256), (2, 2047, 64)):
    bins = torch.rand((n, M + 1), generator=g, device=dev).sort(dim=1).values
    ops.sample_pdf(bins, torch.rand((n, M), generator=g, device=dev), torch.rand((n, K), generator=g, device=dev), want_inds=True)
torch.cuda.synchronize()
print("sanitize_c5: done")
