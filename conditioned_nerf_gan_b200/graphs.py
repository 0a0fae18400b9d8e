"""CUDA-graph replay of the forward for launch-bound shapes.

One ``ImplicitGenerator3d.forward`` is ~17 kernel launches; at small shapes (BASELINE config 1: 64x64, 12+12
samples, batch 1) the kernels take ~0.1 ms and the Python / launch overhead ~0.3 ms.  ``GraphedRender`` captures
the whole forward (RNG draws included: torch's CUDA generator is graph-safe) into one CUDA graph with static
input / output buffers and replays it.
"""
from __future__ import annotations

from typing import Tuple

import torch


class GraphedRender:
    """``render = GraphedRender(generator, z, cam2worlds, **metadata)`` captures; ``pixels, depth = render(z, cam2worlds)``
    copies the new inputs into the static buffers and replays.  Shapes and metadata are fixed at capture time; the
    returned tensors are the static output buffers (valid until the next call)."""

    def __init__(self, generator, z, cam2worlds: torch.Tensor, warmup: int = 3, **metadata):
        self.generator = generator
        self.metadata = dict(metadata)
        self._tuple = isinstance(z, (tuple, list))
        self._z = tuple(t.detach().clone() for t in z) if self._tuple else z.detach().clone()
        self._cam = cam2worlds.detach().clone()
        dev = cam2worlds.device
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):                       # one-time work (function attributes, table caches, allocator pools)
                generator(self._z, self._cam, **self.metadata)
        torch.cuda.current_stream(dev).wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self._pixels, self._depth = generator(self._z, self._cam, **self.metadata)

    def __call__(self, z, cam2worlds: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        if self._tuple:
            for dst, src in zip(self._z, z):
                dst.copy_(src, non_blocking=True)
        else:
            self._z.copy_(z, non_blocking=True)
        self._cam.copy_(cam2worlds, non_blocking=True)
        self.graph.replay()
        return self._pixels, self._depth
