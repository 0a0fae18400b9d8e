"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: pose sharding + frame gather and the
bucketed gradient all-reduce.  The render itself is replaced by a deterministic stub: the N>1 path has
no collective inside the render, so what needs testing on CPU is the sharding arithmetic and the plumbing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conditioned_nerf_gan_b200 import parallel
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


class _StubGenerator:
    """staged_forward that encodes (pose id, fov) into the frame so the gather order can be checked."""

    def staged_forward(self, z, cam2worlds, fov, max_batch_size=8, **meta):
        P, img = cam2worlds.shape[0], meta["img_size"]
        fovs = list(fov) if hasattr(fov, "__len__") else [fov] * P
        assert len(fovs) == P
        ids = cam2worlds[:, 0, 3]
        pixels = ids.view(P, 1, 1, 1).expand(P, 3, img, img).clone() + z[1].sum()
        depth = torch.tensor(fovs, dtype=torch.float32).view(P, 1, 1).expand(P, img, img).clone()
        return pixels, depth


def _worker(rank, world, port, P):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    r, w, _ = parallel.init_distributed("gloo")
    assert (r, w) == (rank, world)
    # ---- pose sharding + ragged gather (BASELINE config 4 plumbing)
    poses = torch.eye(4).repeat(P, 1, 1)
    poses[:, 0, 3] = torch.arange(P, dtype=torch.float32)
    z = (torch.zeros(1, 4, 2, 2, 2), torch.full((1, 8), 0.5 if rank == 0 else -7.0))
    parallel.broadcast_z(z, src=0)
    assert float(z[1][0, 0]) == 0.5
    fov = [30.0 + i for i in range(P)]
    pixels, depth = parallel.render_poses_sharded(_StubGenerator(), z, poses, img_size=4, fov=fov, num_steps=6)
    assert pixels.shape == (P, 3, 4, 4) and depth.shape == (P, 4, 4)
    assert torch.equal(pixels[:, 0, 0, 0], torch.arange(P, dtype=torch.float32) + 4.0)
    assert torch.equal(depth[:, 0, 0], torch.tensor(fov))
    local_p, _ = parallel.render_poses_sharded(_StubGenerator(), z, poses, img_size=4, fov=45.0, num_steps=6, gather=False)
    lo, hi = parallel.shard_range(P, rank, world)
    assert local_p.shape[0] == hi - lo
    # ---- gradient all-reduce of the real generator module (parameters only; no kernel runs on CPU)
    torch.manual_seed(0)
    gen = ImplicitGenerator3d("DOUBLESIREN_FG", 256, 32, 4, 256)
    for i, p in enumerate(gen.parameters()):
        p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
    unused = gen.siren.final_layer.bias
    if rank == 1:
        unused.grad = None                       # find_unused_parameters semantics: missing grad counts as zeros
    n = parallel.allreduce_gradients(gen.parameters(), bucket_bytes=1 << 20)
    assert n >= 2                                # several buckets at 1 MB
    for i, p in enumerate(gen.parameters()):
        expect = 1.5 * (i + 1) if p is not unused else 0.5 * (i + 1)
        assert torch.allclose(p.grad, torch.full_like(p, expect)), (i, float(p.grad.flatten()[0]), expect)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [7, 2, 1])
def test_world2_gloo(P):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, P), nprocs=2, join=True)


def _train_worker(rank, world, port, out):
    """The GAN train step under DDP (gloo, CPU): each rank trains on its shard of the batch; gradients are averaged once per
    optimizer step (micro-batches under no_sync), so both ranks hold identical parameters afterwards and they equal a
    single-process step on the whole batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    parallel.init_distributed("gloo")
    from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator
    from conditioned_nerf_gan_b200.generators.unet3d import UNet3D
    from conditioned_nerf_gan_b200.training import GanTrainStep
    from oracle import nerf_path as oracle
    from oracle import train_step as ts
    torch.set_num_threads(2)
    enc, disc = UNet3D(**ts.TINY_UNET), ProgressiveDiscriminator()
    ts.fill_params(enc, 1)
    ts.fill_params(disc, 2)
    gen = ts.OracleGenerator(ts.TINY_SIREN, oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0))
    sample = ts.tiny_sample()
    draws = ts.tiny_draws()
    lo, hi = parallel.shard_range(ts.TINY_BATCH, rank, world)
    R = ts.tiny_config()["img_size"] ** 2
    shard = {k: v[lo:hi] for k, v in sample.items()}
    shard_draws = {k: (v[lo * R:hi * R] if k == "u_resample" else v[lo:hi]) for k, v in draws.items()}
    md = dict(ts.tiny_config(), draws=shard_draws, r1_lambda=0)      # R1 is a per-rank mean of per-sample terms: same average either way
    tr = GanTrainStep(gen, enc, disc, md, "cpu", amp=False, ddp=True)
    tr.alpha = 0.3
    tr.train_discriminator(shard)
    tr.train_generator(shard)
    flat = torch.cat([p.detach().reshape(-1) for m in (gen, enc, disc) for p in m.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    assert torch.equal(gathered[0], gathered[1]), "ranks diverged after one DDP step"
    if rank == 0:
        torch.save({"norm_G": float(tr.grad_norms["G"]), "norm_E": float(tr.grad_norms["E"]), "norm_D": float(tr.grad_norms["D"]),
                    "g_loss": float(tr.losses["g_loss"]), "checksum": float(flat.double().abs().sum())}, out)
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_gan_train_step(tmp_path):
    from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator
    from conditioned_nerf_gan_b200.generators.unet3d import UNet3D
    from conditioned_nerf_gan_b200.training import GanTrainStep
    from oracle import nerf_path as oracle
    from oracle import train_step as ts
    out = str(tmp_path / "ddp.pt")
    mp.spawn(_train_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    ddp = torch.load(out)
    # single process, whole batch
    enc, disc = UNet3D(**ts.TINY_UNET), ProgressiveDiscriminator()
    ts.fill_params(enc, 1)
    ts.fill_params(disc, 2)
    gen = ts.OracleGenerator(ts.TINY_SIREN, oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0))
    tr = GanTrainStep(gen, enc, disc, dict(ts.tiny_config(), draws=ts.tiny_draws(), r1_lambda=0), "cpu", amp=False)
    tr.alpha = 0.3
    sample = ts.tiny_sample()
    tr.train_discriminator(sample)
    tr.train_generator(sample)
    flat = torch.cat([p.detach().reshape(-1) for m in (gen, enc, disc) for p in m.parameters()])
    # the generator / encoder see the mean over the batch of per-image losses: DDP's average of per-rank means is the same number
    for k in ("norm_G", "norm_E", "norm_D"):
        assert abs(ddp[k] - float(tr.grad_norms[k[-1]])) <= 2e-3 * abs(ddp[k]) + 1e-6, (k, ddp[k], float(tr.grad_norms[k[-1]]))
    assert abs(ddp["checksum"] - float(flat.double().abs().sum())) <= 1e-5 * ddp["checksum"]


def test_shard_range_partitions():
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_world1_is_noop():
    assert parallel.allreduce_gradients([torch.nn.Parameter(torch.zeros(3))], world=1) == 0
