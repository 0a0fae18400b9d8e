"""Inference drivers around the generator: the reference's ``generate_img``
(utils.py:60-82) and the camera path + frame loop of ``Inferencer.render_video`` (inference.py:441-486),
without the per-frame Python loop and host sync: frames are rendered by ``staged_forward`` (optionally
sharded over the ranks of the default process group) and come back as one tensor.
Video encoding, checkpoints and datasets stay with the caller (out of scope).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import torch

from . import parallel
from .generators.volumetric_rendering import create_cam2world_matrix


def generate_img(generator, z, cam2worlds: torch.Tensor, metadata: dict) -> Tuple[torch.Tensor, torch.Tensor]:
    """utils.py:60-82: no-grad render, both results on the host, the depth map stacked to 3 channels."""
    with torch.no_grad():
        img, depth_map = generator(z, cam2worlds, **metadata)
        img = img.cpu()
        depth_map = depth_map.cpu()
        depth_map = torch.stack([depth_map] * 3, 1)
    return img, depth_map


def video_camera_path(num_frames: int, fps: int, cam_r_start: float, cam_r_end: float, up_direction: str = "y", device=None):
    """inference.py:442-477: the camera origins of the turntable / elevation sweep and the 60 -> 30 degree fov ramp.
    Returns (cam2world [F,4,4] on ``device``, fov [F] numpy).  ``num_frames`` must be a multiple of 4 and >= 4 * fps."""
    if num_frames % 4 or num_frames // 4 < fps:
        raise ValueError("num_frames must be a multiple of 4 and at least 4 * fps")
    q, h = num_frames // 4, num_frames // 2
    theta = np.concatenate([np.linspace(1e-5, np.pi / 2 - 1e-5, h), np.linspace(np.pi / 2 - 1e-5, 1e-5, q),
                            np.linspace(1e-5, np.pi / 4 - 1e-5, fps), np.asarray([np.pi / 4 - 1e-5] * (q - fps))], axis=0)
    phi = np.concatenate([np.linspace(0, np.pi * 2, h), np.linspace(np.pi * 2, np.pi * 5 / 4, fps),
                          np.asarray([np.pi * 5 / 4] * (q - fps)), np.linspace(np.pi * 5 / 4, 0, q)], axis=0)
    r = np.linspace(cam_r_start, cam_r_end, num_frames)
    fov = np.linspace(60, 30, num_frames)
    origin = np.zeros((num_frames, 3))
    origin[:, 0] = r * np.sin(theta) * np.cos(phi)
    side, top = r * np.sin(theta) * np.sin(phi), r * np.cos(theta)
    if up_direction == "z":
        origin[:, 1], origin[:, 2] = side, top
    elif up_direction == "y":
        origin[:, 2], origin[:, 1] = side, top
    else:
        raise ValueError("up_direction must be 'y' or 'z'")
    origin_t = torch.from_numpy(origin).type(torch.float32).to(device)
    return create_cam2world_matrix(origin_t, up_direction, device), fov


@torch.no_grad()
def render_video_frames(generator, z, metadata: dict, num_frames: int = 64, fps: int = 8, up_direction: str = "y",
                        max_batch_size: int = 4, sharded: bool = False) -> torch.Tensor:
    """The frame loop of inference.py:478-486 in one call: [num_frames, 3, H, W] frames on the host.

    ``metadata`` is the curriculum dict (``cam_r_start`` / ``cam_r_end`` give the radius ramp, ``fov`` is replaced by
    the per-frame ramp, ``nerf_noise`` is forced to 0).  With ``sharded=True`` the poses are split over the ranks of the
    default process group (``parallel.render_poses_sharded``) and every rank returns all frames."""
    vol = z[0] if isinstance(z, (tuple, list)) else z
    cam2world, fov = video_camera_path(num_frames, fps, metadata["cam_r_start"], metadata["cam_r_end"], up_direction, vol.device)
    meta = {k: v for k, v in metadata.items() if k != "fov"}
    if sharded:
        pixels, _ = parallel.render_poses_sharded(generator, z, cam2world, fov=list(fov), max_batch_size=max_batch_size, **meta)
    else:
        pixels, _ = generator.staged_forward(z, cam2world, fov=list(fov), max_batch_size=max_batch_size, **meta)
    return pixels.cpu()
