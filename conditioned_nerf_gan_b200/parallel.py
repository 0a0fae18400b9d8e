"""Multi-GPU plumbing of the rendering path: one process per GPU, ``torch.distributed``.

The path shards by independent units: every ray / image / pose is independent
given (feature volume, global feature, weights, cam2world), so there is NO collective inside the
render.  What crosses NVLink:

* training (reference: DDP over gloo, utils.py:322-326, train.py:40): only the gradient
  all-reduce of the generator parameters -- ``torch.nn.parallel.DistributedDataParallel`` over NCCL,
  or ``allreduce_gradients`` when the caller accumulates micro-batches itself (the reference fires
  one all-reduce per micro-batch backward, utils.py:711; this does one per optimizer step);
* inference (reference: per-frame loop, inference.py:478-486): poses are sharded, each rank
  renders its slice with ``staged_forward``, frames are all-gathered (``render_poses_sharded``).

All helpers run on the gloo backend too (CPU tests with world_size 2).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Optional, Tuple

import torch
import torch.distributed as dist


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Initialise the default process group from the torchrun environment (RANK / WORLD_SIZE /
    LOCAL_RANK / MASTER_ADDR / MASTER_PORT).  Returns (rank, world, local_rank); a no-op for world 1."""
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def bind_to_gpu_numa_node(local_rank: int) -> Optional[List[int]]:
    """Pin this process to the CPU cores NVML reports as closest to GPU ``local_rank`` (its NUMA node), so that pinned host
    buffers allocated afterwards land in that node's memory and the H2D / D2H copies of the ranks of one box do not all cross
    the same socket link.  Returns the core list, or None when NVML / affinity control is unavailable (nothing changes then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (os.cpu_count() + 63) // 64)
        cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [start, stop) of n units for ``rank`` (the first n % world ranks get one more)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def allreduce_gradients(params: Iterable[torch.nn.Parameter], world: Optional[int] = None, bucket_bytes: int = 64 << 20) -> int:
    """Average ``.grad`` of the given parameters over the default group in flat buckets (one NCCL
    all-reduce per ~64 MB; the generator's 1.5 M parameters are a single bucket).  Parameters without
    a gradient contribute zeros (DDP's ``find_unused_parameters=True`` semantics, utils.py:325).
    Returns the number of collectives issued."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return 0
    params = [p for p in params if p.requires_grad]
    n_coll = 0
    bucket: List[torch.nn.Parameter] = []
    size = 0

    def flush():
        nonlocal bucket, size, n_coll
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).float() for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for p in bucket:
            g = flat[off:off + p.numel()].view_as(p).to(p.dtype)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += p.numel()
        n_coll += 1
        bucket, size = [], 0

    for p in params:
        bucket.append(p)
        size += p.numel() * 4
        if size >= bucket_bytes:
            flush()
    flush()
    return n_coll


@torch.no_grad()
def render_poses_sharded(generator, z, cam2worlds: torch.Tensor, *, max_batch_size: int = 8, gather: bool = True, **metadata):
    """Render P poses of ONE object across the ranks of the default group (BASELINE config 4).

    Every rank holds the same ``z`` (broadcast it first if only rank 0 ran the encoder) and the full
    ``cam2worlds`` [P,4,4]; rank r renders poses ``shard_range(P, r, world)`` with
    ``generator.staged_forward`` and, if ``gather``, receives all frames: pixels [P,3,H,W], depth [P,H,W].
    ``fov`` may be a per-pose sequence.  No collective runs inside the render."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    P = cam2worlds.shape[0]
    lo, hi = shard_range(P, rank, world)
    meta = dict(metadata)
    fov = meta.pop("fov")
    if hasattr(fov, "__len__"):
        fov = list(fov)[lo:hi]
    img = meta["img_size"]
    if hi > lo:
        pixels, depth = generator.staged_forward(z, cam2worlds[lo:hi], fov=fov, max_batch_size=max_batch_size, **meta)
    else:
        pixels = torch.empty((0, 3, img, img), dtype=torch.float32, device=cam2worlds.device)
        depth = torch.empty((0, img, img), dtype=torch.float32, device=cam2worlds.device)
    if world == 1 or not gather:
        return pixels, depth
    # ragged all-gather: pad every shard to the largest one
    most = -(-P // world)
    pad_p = torch.zeros((most, 3, img, img), dtype=pixels.dtype, device=pixels.device)
    pad_d = torch.zeros((most, img, img), dtype=depth.dtype, device=depth.device)
    pad_p[: hi - lo], pad_d[: hi - lo] = pixels, depth
    all_p = [torch.empty_like(pad_p) for _ in range(world)]
    all_d = [torch.empty_like(pad_d) for _ in range(world)]
    dist.all_gather(all_p, pad_p)
    dist.all_gather(all_d, pad_d)
    counts = [shard_range(P, r, world) for r in range(world)]
    return (torch.cat([all_p[r][: b - a] for r, (a, b) in enumerate(counts)]),
            torch.cat([all_d[r][: b - a] for r, (a, b) in enumerate(counts)]))


def broadcast_z(z, src: int = 0):
    """Broadcast the encoder output (feature volume, global feature) from ``src`` (33.5 MB + 1 KB per object)."""
    if dist.is_initialized() and dist.get_world_size() > 1:
        for t in z:
            dist.broadcast(t, src=src)
    return z
