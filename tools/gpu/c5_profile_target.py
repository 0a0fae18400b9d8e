import os, sys, torch
sys.path.insert(0, os.getcwd())
from conditioned_nerf_gan_b200 import ops
dev = "cuda"; n = 1 << 20; S = 64
g = torch.Generator(device=dev).manual_seed(0)
rs = torch.randn((n, S, 4), generator=g, device=dev); rs2 = torch.randn((n, S, 4), generator=g, device=dev)
t = torch.rand((n, S), generator=g, device=dev).mul_(1.7).add_(0.25).sort(dim=1).values
t2 = torch.rand((n, S), generator=g, device=dev).mul_(1.7).add_(0.25)
w = torch.rand((n, S), generator=g, device=dev); u = torch.rand((n, S), generator=g, device=dev)
rays = torch.nn.functional.normalize(torch.randn((1024, 3), generator=g, device=dev), dim=-1)
for _ in range(2):
    ops.merge_composite(rs2, rs, t2, t, None, rays, n // 1024, 32, 32, 0.0, "relu", True, False)
    ops.resample_from_coarse(t, w, u)
torch.cuda.synchronize()
