import os, sys, torch
sys.path.insert(0, os.getcwd())
from conditioned_nerf_gan_b200 import ops
dev="cuda"; n=1<<21
for S in (200, 256):
    g=torch.Generator(device=dev).manual_seed(0)
    t=torch.rand((n,S),generator=g,device=dev).mul_(1.7).add_(0.25).sort(dim=1).values
    w=torch.rand((n,S),generator=g,device=dev); u=torch.rand((n,S),generator=g,device=dev)
    out=ops.resample_from_coarse(t,w,u)
    for _ in range(3): ops.resample_from_coarse(t,w,u)
    torch.cuda.synchronize()
    s,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(5): ops.resample_from_coarse(t,w,u)
    e.record(); torch.cuda.synchronize()
    ms=s.elapsed_time(e)/5
    print(S, f"{ms:.3f} ms", f"{n*16*S/ms/1e6:.0f} GB/s", "checksum", float(out[0].double().sum() if isinstance(out,(tuple,list)) else out.double().sum()))
