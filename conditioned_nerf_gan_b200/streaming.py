"""Host-buffer front end of the renderer: H2D copies, kernels and D2H copies on three CUDA streams.

The reference moves every batch to the device synchronously and reads every image back with ``.cpu()``
(utils.py:625-630, utils.py:79-80, inference.py:478-486).  ``render_host_batches`` keeps the same
per-batch contract -- inputs start in (pinned) host memory, images end in (pinned) host memory -- but
overlaps the copy of batch i+1 and the read-back of batch i-1 with the kernels of batch i.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Optional, Sequence, Tuple

import torch


class _Slot:
    """Device input tensors + pinned host outputs of one in-flight batch, with the events guarding their reuse."""

    def __init__(self):
        self.inputs: Optional[Tuple[torch.Tensor, ...]] = None
        self.h2d_done = torch.cuda.Event()
        self.compute_done = torch.cuda.Event()
        self.d2h_done = torch.cuda.Event()
        self.pixels_h: Optional[torch.Tensor] = None
        self.depth_h: Optional[torch.Tensor] = None
        self.busy = False


@torch.no_grad()
def render_host_batches(generator, batches: Iterable[Sequence[torch.Tensor]], metadata: dict, device=None,
                        depth: int = 3) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    """Render a stream of host batches ``(volume[B,C,D,H,W], global[B,z], cam2world[B,4,4])`` (CPU tensors,
    ideally pinned) through ``generator`` and yield ``(pixels, depth)`` as pinned host tensors, in order.

    ``depth`` batches are in flight (3: one being copied in, one computing, one being read back).  A yielded
    pair stays valid until ``depth - 1`` further results have been taken from the iterator.
    """
    if depth < 2:
        raise ValueError("depth must be >= 2")
    device = torch.device(device if device is not None else "cuda")
    copy_in, copy_out = torch.cuda.Stream(device), torch.cuda.Stream(device)
    compute = torch.cuda.current_stream(device)
    slots = [_Slot() for _ in range(depth)]
    it = iter(batches)

    def stage(slot: _Slot, batch) -> None:
        # persistent device buffers per slot (no allocator traffic in steady state); the kernels that last read them
        # must be done before they are overwritten
        if slot.inputs is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(slot.inputs, batch)):
            slot.inputs = tuple(torch.empty(h.shape, dtype=h.dtype, device=device) for h in batch)
        copy_in.wait_event(slot.compute_done)
        with torch.cuda.stream(copy_in):
            for d, h in zip(slot.inputs, batch):
                d.copy_(h, non_blocking=True)
            slot.h2d_done.record(copy_in)

    def take(slot: _Slot):
        slot.d2h_done.synchronize()
        slot.busy = False
        return slot.pixels_h, slot.depth_h

    nxt = next(it, None)
    if nxt is not None:
        stage(slots[0], nxt)
    i = 0
    while nxt is not None:
        slot = slots[i % depth]
        nxt = next(it, None)
        if nxt is not None:
            nslot = slots[(i + 1) % depth]
            if nslot.busy:                      # batch i+1-depth: hand its image out before its buffers are reused
                yield take(nslot)
            stage(nslot, nxt)
        compute.wait_event(slot.h2d_done)
        vol, glob, cam = slot.inputs
        pixels, dep = generator((vol, glob), cam, **metadata)
        slot.compute_done.record(compute)
        copy_out.wait_event(slot.compute_done)
        with torch.cuda.stream(copy_out):
            if slot.pixels_h is None or slot.pixels_h.shape != pixels.shape:
                slot.pixels_h = torch.empty(pixels.shape, dtype=pixels.dtype).pin_memory()
                slot.depth_h = torch.empty(dep.shape, dtype=dep.dtype).pin_memory()
            slot.pixels_h.copy_(pixels, non_blocking=True)
            slot.depth_h.copy_(dep, non_blocking=True)
            pixels.record_stream(copy_out)
            dep.record_stream(copy_out)
            slot.d2h_done.record(copy_out)
        slot.busy = True
        i += 1
    # drain, oldest first
    for k in range(depth):
        s = slots[(i + k) % depth]
        if s.busy:
            yield take(s)
