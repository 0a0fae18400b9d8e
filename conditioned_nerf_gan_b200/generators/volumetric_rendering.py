"""Host mirror of the reference's rendering functions (generators/volumetric_rendering.py).

Same names, argument meaning and error behaviour as the reference; the hot ones launch the
sm_100a kernels of libcng_b200 (K3 ``cng_composite_fwd``, K4 ``cng_sample_pdf``).  Ray
generation, jitter and the camera->world transform have no standalone kernel: inside the
generator they are fused into K1 (``ops.raymarch_gather_coarse``); the functions of those names
here only serve the reference's secondary callers (feature_volume/voxel2img.py) and work on the
small per-image tables.

Camera helpers (``sample_camera_positions``, ``create_cam2world_matrix``) produce B x 4 x 4
matrices on the host side exactly like the reference; they are inputs of the path, not part of it.
"""
from __future__ import annotations

import math
from functools import lru_cache
from typing import Tuple

import numpy as np
import torch

from .. import ops

__all__ = ["fancy_integration", "get_initial_rays_trig", "perturb_points", "transform_sampled_points",
           "sample_camera_positions", "create_cam2world_matrix", "sample_pdf", "distance2depth",
           "normalize_vecs", "camera_tables"]


def normalize_vecs(vectors: torch.Tensor) -> torch.Tensor:
    """generators/math_utils_torch.py:16-20."""
    return vectors / torch.norm(vectors, dim=-1, keepdim=True)


@lru_cache(maxsize=64)
def _camera_tables_cpu(W: int, H: int, S: int, fov: float, ray_start: float, ray_end: float):
    # Built with the reference's own op sequence on the CPU (volumetric_rendering.py:77-93) so the
    # tables agree with it to the last bit; they are tiny (R*3 + S floats) and cached per shape.
    x, y = torch.meshgrid(torch.linspace(-1, 1, W), torch.linspace(-1, 1, H), indexing="ij")
    x = x.T.flatten()
    y = y.T.flatten()
    z = torch.ones_like(x) / np.tan((2 * math.pi * fov / 360) / 2)
    rays = normalize_vecs(torch.stack([x, y, z], -1)).contiguous()
    t_lin = torch.linspace(ray_start, ray_end, S).contiguous()
    return rays, t_lin


_device_tables = {}


def camera_tables(resolution, num_steps, fov, ray_start, ray_end, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """rays_d_cam [R,3] (ray p = row*W + col) and the coarse distances t_lin [S], on ``device``."""
    W, H = resolution
    key = (int(W), int(H), int(num_steps), float(fov), float(ray_start), float(ray_end), str(device))
    hit = _device_tables.get(key)
    if hit is None:
        rays, t_lin = _camera_tables_cpu(*key[:-1])
        hit = (rays.to(device), t_lin.to(device))
        if len(_device_tables) > 256:
            _device_tables.clear()
        _device_tables[key] = hit
    return hit


def get_initial_rays_trig(n, num_steps, device, fov, resolution, ray_start, ray_end):
    """volumetric_rendering.py:73-100.  Returns points [n,R,S,3], z_vals [n,R,S,1], rays_d_cam [n,R,3]
    (broadcast views of the per-image tables; the generator itself never materialises them)."""
    rays, t_lin = camera_tables(resolution, num_steps, fov, ray_start, ray_end, device)
    R = rays.shape[0]
    z_vals = t_lin.reshape(1, num_steps, 1).expand(R, num_steps, 1)
    points = rays.unsqueeze(1) * z_vals
    return (points.unsqueeze(0).expand(n, -1, -1, -1), z_vals.unsqueeze(0).expand(n, -1, -1, -1),
            rays.unsqueeze(0).expand(n, -1, -1))


def perturb_points(points, z_vals, ray_directions, device):
    """volumetric_rendering.py:103-110 (compatibility entry; fused into K1 on the product path)."""
    spacing = z_vals[:, :, 1:2, :] - z_vals[:, :, 0:1, :]
    offset = (torch.rand(z_vals.shape, device=device) - 0.5) * spacing
    return points + offset * ray_directions.unsqueeze(2), z_vals + offset


def transform_sampled_points(points, z_vals, ray_directions, device, cam2worlds):
    """volumetric_rendering.py:113-199 (compatibility entry; fused into K1 on the product path)."""
    n, R, S, _ = points.shape
    points, z_vals = perturb_points(points, z_vals, ray_directions, device)
    rot, trans = cam2worlds[:, :3, :3], cam2worlds[:, :3, 3]
    pts_w = torch.einsum("bij,brsj->brsi", rot, points) + trans[:, None, None, :]
    dirs_w = torch.einsum("bij,brj->bri", rot, ray_directions)
    origins = trans[:, None, :].expand(n, R, 3)
    return pts_w, z_vals, dirs_w, origins


def fancy_integration(rgb_sigma, z_vals, device, noise_std=0.5, last_back=False, white_back=False,
                      clamp_mode=None, fill_mode=None, noise=None):
    """volumetric_rendering.py:18-70 on K3 (one warp per ray).

    rgb_sigma [B,R,S,4], z_vals [B,R,S,1] -> rgb [B,R,3], depth [B,R,1], weights [B,R,S,1].
    ``noise`` replaces the reference's ``torch.randn(sigmas.shape)`` draw (made here, on the
    device, when not given -- also when noise_std == 0, which keeps the RNG stream aligned).
    """
    if fill_mode is not None:
        raise NotImplementedError("fill_mode is a debug option no caller of the reference uses")
    ops.clamp_code(clamp_mode)    # TypeError("Need to choose clamp mode") before anything is drawn
    if noise is None:
        noise = torch.randn(z_vals.shape, device=rgb_sigma.device)
    if torch.is_grad_enabled() and rgb_sigma.requires_grad:
        rgb, dist, w = _CompositeFn.apply(rgb_sigma, z_vals, noise, noise_std, clamp_mode, white_back, last_back)
    else:
        rgb, dist, w = ops.composite_fwd(rgb_sigma, z_vals, noise, noise_std, clamp_mode, white_back, last_back)
    return rgb, dist.unsqueeze(-1), w.unsqueeze(-1)


class _CompositeFn(torch.autograd.Function):
    """fancy_integration with gradients to rgb_sigma (cng_composite_bwd); the returned weights are not differentiable
    (the generator only uses them under no_grad, generators.py:111-121)."""

    @staticmethod
    def forward(ctx, rgb_sigma, z_vals, noise, noise_std, clamp_mode, white_back, last_back):
        rgb, dist, w = ops.composite_fwd(rgb_sigma, z_vals, noise, noise_std, clamp_mode, white_back, last_back)
        ctx.save_for_backward(rgb_sigma, z_vals, noise)
        ctx.cfg = (noise_std, clamp_mode, white_back, last_back)
        ctx.mark_non_differentiable(w)
        return rgb, dist, w

    @staticmethod
    def backward(ctx, d_rgb, d_dist, _d_w):
        rgb_sigma, z_vals, noise = ctx.saved_tensors
        noise_std, clamp_mode, white_back, last_back = ctx.cfg
        d = ops.composite_bwd(rgb_sigma, z_vals, noise, d_rgb, d_dist, noise_std, clamp_mode, white_back, last_back)
        return d, None, None, None, None, None, None


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5, u=None, return_inds=False):
    """volumetric_rendering.py:297-342 on K4.  bins [n,M+1], weights [n,M] -> samples [n,N_importance].

    ``u`` replaces the reference's ``torch.rand(N_rays, N_importance)`` (or the linspace of det=True).
    """
    n = weights.shape[0]
    if u is None:
        if det:
            u = torch.linspace(0, 1, N_importance, device=bins.device).expand(n, N_importance)
        else:
            u = torch.rand(n, N_importance, device=bins.device)
    return ops.sample_pdf(bins, weights, u, eps, want_inds=return_inds)


def distance2depth(distance: torch.Tensor, ray: torch.Tensor) -> torch.Tensor:
    """volumetric_rendering.py:345-356."""
    return ray[..., -1:] * distance


def sample_camera_positions(device, up_direction, cam_r_start=0, cam_r_end=1, n=1):
    """volumetric_rendering.py:212-238: n random camera origins on a spherical shell (numpy RNG)."""
    assert up_direction in ["y", "z"]
    theta = np.clip(np.arccos(1 - np.random.rand(n)), 1e-5, np.pi - 1e-5)
    phi = np.random.rand(n) * np.pi * 2
    r = np.random.rand(n) * (cam_r_end - cam_r_start) + cam_r_start
    origin = np.zeros((n, 3))
    origin[:, 0] = r * np.sin(theta) * np.cos(phi)
    side, top = r * np.sin(theta) * np.sin(phi), r * np.cos(theta)
    if up_direction == "z":
        origin[:, 1], origin[:, 2] = side, top
    else:
        origin[:, 2], origin[:, 1] = side, top
    return torch.from_numpy(origin).type(torch.float32).to(device)


def create_cam2world_matrix(origin, up_direction, device=None):
    """volumetric_rendering.py:255-287: look-at matrix towards the world origin."""
    assert up_direction in ["y", "z"]
    fwd = normalize_vecs(-origin)
    axis = [0, 1, 0] if up_direction == "y" else [0, 0, 1]
    up = torch.tensor(axis, dtype=torch.float, device=device).expand_as(fwd)
    left = normalize_vecs(torch.cross(up, fwd, dim=-1))
    up = normalize_vecs(torch.cross(fwd, left, dim=-1))
    n = fwd.shape[0]
    rot = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, -up, fwd), dim=-1)
    trans = torch.eye(4, device=device).unsqueeze(0).repeat(n, 1, 1)
    trans[:, :3, 3] = origin
    return trans @ rot
