"""Dense density-grid extraction through the SIREN-only boundary: mirror of the reference's
``extract_shapes.create_samples`` / ``sample_generator`` (extract_shapes.py:15-78).

Same sample order and coordinates as the reference -- including its quirk that the x / y grid indices
are computed with a float division and are therefore not integers (extract_shapes.py:26-27) -- but the
coordinates are generated on the device chunk by chunk instead of as one 16.7 M x 3 host tensor, and each
chunk is one ``generator.siren`` call (``cng_gather_points`` + ``cng_film_siren_fwd``).
"""
from __future__ import annotations

from typing import Sequence, Tuple

import numpy as np
import torch


def _coords(index: torch.Tensor, N: int, corner, voxel_size: float) -> torch.Tensor:
    """extract_shapes.py:24-33 for the given flat indices (int64): returns [P, 3] fp32."""
    f = index.float()
    out = torch.empty((index.numel(), 3), dtype=torch.float32, device=index.device)
    out[:, 2] = (index % N).float()
    out[:, 1] = (f / N) % N
    out[:, 0] = ((f / N) / N) % N
    out[:, 0] = (out[:, 0] * voxel_size) + corner[2]
    out[:, 1] = (out[:, 1] * voxel_size) + corner[1]
    out[:, 2] = (out[:, 2] * voxel_size) + corner[0]
    return out


def create_samples(N: int = 256, voxel_origin: Sequence[float] = (0, 0, 0), cube_length: float = 2.0, device="cpu"):
    """extract_shapes.py:15-37.  Returns (samples [1, N^3, 3], corner, voxel_size); ``voxel_origin`` is the cube centre."""
    corner = np.array(voxel_origin) - cube_length / 2
    voxel_size = cube_length / (N - 1)
    index = torch.arange(0, N ** 3, dtype=torch.int64, device=device)
    return _coords(index, N, corner, voxel_size).unsqueeze(0), corner, voxel_size


@torch.no_grad()
def sample_generator(generator, z, voxel_resolution: int = 256, voxel_origin: Sequence[float] = (0, 0, 0),
                     cube_length: float = 1.2, psi: float = 0.5, max_points: int = 1 << 21) -> np.ndarray:
    """extract_shapes.py:40-78: the raw sigma channel of ``generator.siren`` on a dense ``voxel_resolution``^3 grid, as a
    numpy array [N, N, N].  ``psi`` is accepted and unused, as in the reference."""
    N = int(voxel_resolution)
    corner = np.array(voxel_origin) - cube_length / 2
    voxel_size = cube_length / (N - 1)
    vol = z[0] if isinstance(z, (tuple, list)) else z
    dev = vol.device
    sigmas = torch.empty((N ** 3,), dtype=torch.float32, device=dev)
    for head in range(0, N ** 3, max_points):
        stop = min(N ** 3, head + max_points)
        pts = _coords(torch.arange(head, stop, dtype=torch.int64, device=dev), N, corner, voxel_size).unsqueeze(0)
        sigmas[head:stop] = generator.siren(pts, z, N, 1)[0, :, 3]
    return sigmas.reshape(N, N, N).cpu().numpy()
