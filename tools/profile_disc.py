import os, sys
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
from torch.profiler import profile, ProfilerActivity
from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator
dev = torch.device("cuda"); b = 8
real = torch.rand((b, 3, 128, 128), device=dev) * 2 - 1
disc = ProgressiveDiscriminator().to(dev)
def step():
    r = real.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        rp = disc(r, 1.0)
    g = torch.autograd.grad(rp.sum() * 1024.0, r, create_graph=True)[0] / 1024.0
    with torch.autocast("cuda", dtype=torch.float16):
        loss = F.softplus(-rp).mean() + 5.0 * (g.reshape(b, -1).norm(2, dim=1) ** 2).mean()
    (loss * 1024.0).backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages(group_by_input_shape=True).table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=50, max_shapes_column_width=110))
