"""``ImplicitGenerator3d``: drop-in for the reference generator (generators/generators.py:9-197).

Same constructor, ``forward(z, cam2worlds, **curriculum_metadata) -> (pixels[B,3,H,W],
depth[B,H,W])``, ``set_device``, ``epoch`` / ``step`` attributes and ``state_dict`` keys.  The body
is five kernel launches per SIREN pass instead of ~100 eager ATen kernels:

    K1  cng_raymarch_gather_coarse   rays + jitter + cam->world + trilinear gather   (a1-a4)
    K2  cng_film_siren_fwd           all FiLM layers + head, tcgen05 / TMEM          (a5-a7)
    K3  cng_composite_fwd            coarse weights                                  (a8)
    K4  cng_resample_from_coarse     inverse-CDF resampling                          (a9)
    K1' cng_raymarch_gather_fine     o + d*t + gather                                (a10, a4)
    K2  cng_film_siren_fwd           fine pass
    K3' cng_merge_composite          merge by depth + composite + NCHW image + depth (a11, a8, a12)

sequenced by ONE C-ABI call, ``cng_render_fwd`` (the stage-by-stage path remains for the taps the tests compare).

Random draws are made with ``torch.rand`` / ``torch.randn`` on the device in the reference's order
and shapes (rand[B,R,S,1], randn[B,R,S,1], rand[B*R,S], randn[B,R,2S,1]), or taken
from the optional ``draws`` keyword (a dict with keys u_jitter, noise_coarse, u_resample,
noise_final) so that tests can replay the oracle's numbers.

``staged_forward`` (not in the reference; the name comes from upstream pi-GAN) is the chunked
no-grad inference entry used for video rendering.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn as nn

from .. import ops
from . import siren
from .volumetric_rendering import camera_tables


class ImplicitGenerator3d(nn.Module):
    def __init__(self, siren_type, z_dim, input_dim, output_dim, hidden_dim, drop_out=0):
        super().__init__()
        self.z_dim = z_dim
        SIREN = getattr(siren, siren_type)          # unknown name -> AttributeError, as generators.py:15
        self.siren = SIREN(z_dim=z_dim, input_dim=input_dim, output_dim=output_dim, hidden_dim=hidden_dim,
                           drop_out=drop_out, device=None)
        self.epoch = 0
        self.step = 0
        self.device = None

    def set_device(self, device):
        self.device = device
        self.siren.device = device

    # ------------------------------------------------------------------------------------------
    def forward(self, z, cam2worlds, img_size, fov, ray_start, ray_end, num_steps, hierarchical_sample, **kwargs):
        """generators/generators.py:33-187.  ``kwargs`` is the curriculum metadata dict: ``clamp_mode``
        and ``nerf_noise`` are required (KeyError otherwise, as in the reference), ``white_back`` /
        ``last_back`` default to False, every other key is ignored."""
        volume, global_feature = self.siren.split_z(z)
        if getattr(self.siren, "library_mlp", False):
            # TALLSIREN / TALLSIREN_dgx / SHORTSIREN_FG_Pyrmd: the library's kernels around a PyTorch MLP (generators/siren_library.py)
            from .autograd import render_library
            return render_library(self, z, cam2worlds, img_size, fov, ray_start, ray_end, num_steps, hierarchical_sample, kwargs)
        needs_grad = torch.is_grad_enabled() and (
            (volume is not None and volume.requires_grad) or (global_feature is not None and global_feature.requires_grad)
            or any(p.requires_grad for p in self.siren.parameters()))
        if needs_grad:
            from .autograd import render_with_grad
            return render_with_grad(self, volume, global_feature, cam2worlds, img_size, fov, ray_start, ray_end,
                                    num_steps, hierarchical_sample, kwargs)
        out = self._render(volume, global_feature, cam2worlds, img_size, fov, ray_start, ray_end, num_steps,
                           hierarchical_sample, kwargs)
        return out["pixels"], out["depth"]

    @torch.no_grad()
    def _render(self, volume, global_feature, cam2worlds, img_size, fov, ray_start, ray_end, num_steps,
                hierarchical_sample, kwargs, taps: bool = False, vol_cl: Optional[torch.Tensor] = None,
                film=None) -> Dict[str, torch.Tensor]:
        clamp_mode, nerf_noise = kwargs["clamp_mode"], kwargs["nerf_noise"]
        white_back, last_back = kwargs.get("white_back", False), kwargs.get("last_back", False)
        ops.clamp_code(clamp_mode)
        draws = kwargs.get("draws") or {}
        B = cam2worlds.shape[0]
        S, R = int(num_steps), int(img_size) * int(img_size)
        dev = cam2worlds.device
        net = self.siren
        rays_d_cam, t_lin = camera_tables((img_size, img_size), S, fov, ray_start, ray_end, dev)
        freq, phase = film if film is not None else net.film_parameters(global_feature, B, dev)
        if net.latent:
            return self._render_latent(net, freq, phase, cam2worlds, rays_d_cam, t_lin, img_size, S, hierarchical_sample, kwargs, taps)
        if vol_cl is None:
            vol_cl = ops.volume_to_channels_last(volume)
        C = vol_cl.shape[-1]
        out: Dict[str, torch.Tensor] = {}

        def draw(name, fn, shape):
            t = draws.get(name)
            return fn(shape, device=dev) if t is None else t.to(dev)

        net.check_dropout()
        if not taps and kwargs.get("fused_call", True) and not net.res_add_mask:
            # production path: one C-ABI call (cng_render_fwd) sequences K1, K2, K3, K4, K1', K2, K3'; draws in the reference's order
            u_jitter = draw("u_jitter", torch.rand, (B, R, S, 1))
            ws, bs = net.layer_parameters()
            if hierarchical_sample:
                noise_c = draw("noise_coarse", torch.randn, (B, R, S, 1))
                u_re = draw("u_resample", torch.rand, (B * R, S))
                noise_f = draw("noise_final", torch.randn, (B, R, 2 * S, 1))
            else:
                noise_c, u_re = None, None
                noise_f = draw("noise_final" if "noise_final" in draws else "noise_coarse", torch.randn, (B, R, S, 1))
            out["pixels"], out["depth"] = ops.render_fwd(vol_cl, cam2worlds, rays_d_cam, t_lin, ws, bs, freq, phase, net.final_layer.weight,
                                                         net.final_layer.bias, net.sigmoid_rgb, net.precision, u_jitter, noise_c, u_re,
                                                         noise_f, img_size, img_size, hierarchical_sample, nerf_noise, clamp_mode,
                                                         white_back, last_back)
            return out

        u_jitter = draw("u_jitter", torch.rand, (B, R, S, 1))
        feat_c, t_c, pts_c = ops.raymarch_gather_coarse(vol_cl, cam2worlds, rays_d_cam, t_lin, u_jitter, img_size,
                                                        img_size, want_points=taps)
        coarse = net.mlp(feat_c.view(B, R * S, C), freq, phase)
        if hierarchical_sample:
            noise_c = draw("noise_coarse", torch.randn, (B, R, S, 1))
            _, _, w_c = ops.composite_fwd(coarse.view(B, R, S, 4), t_c, noise_c, nerf_noise, clamp_mode)
            u_re = draw("u_resample", torch.rand, (B * R, S))
            t_f, inds = ops.resample_from_coarse(t_c, w_c, u_re, want_inds=True) if taps else (
                ops.resample_from_coarse(t_c, w_c, u_re), None)
            feat_f, pts_f = ops.raymarch_gather_fine(vol_cl, cam2worlds, rays_d_cam, t_f, img_size, img_size,
                                                     want_points=taps)
            fine = net.mlp(feat_f.view(B, R * S, C), freq, phase)
            noise_f = draw("noise_final", torch.randn, (B, R, 2 * S, 1))
            res = ops.merge_composite(fine, coarse, t_f, t_c, noise_f, rays_d_cam, B, img_size, img_size, nerf_noise,
                                      clamp_mode, white_back, last_back, taps=taps)
            if taps:
                out.update(weights_coarse=w_c, t_fine=t_f.view(B, R, S), resample_inds=inds, points_fine=pts_f,
                           rgb_sigma_fine=fine.view(B, R, S, 4), feat_fine=feat_f)
        else:
            # generators.py:172-180: the only composite draws one randn of the coarse shape
            noise_f = draw("noise_final" if "noise_final" in draws else "noise_coarse", torch.randn, (B, R, S, 1))
            res = ops.merge_composite(None, coarse, None, t_c, noise_f, rays_d_cam, B, img_size, img_size, nerf_noise,
                                      clamp_mode, white_back, last_back, taps=taps)
        out["pixels"], out["depth"] = res[0], res[1]
        if taps:
            out.update(points_coarse=pts_c, t_coarse=t_c, rgb_sigma_coarse=coarse.view(B, R, S, 4), feat_coarse=feat_c,
                       rgb=res[2]["rgb"], dist=res[2]["dist"], merge_order=res[2]["order"])
        return out

    @torch.no_grad()
    def _render_latent(self, net, freq, phase, cam2worlds, rays_d_cam, t_lin, img_size, S, hierarchical_sample, kwargs, taps):
        """The same sequence for the position-input SIREN (``SHORTSIREN``, siren.py:1172-1224): K1 in points-only mode (no volume,
        no gather), the positions as the MLP's operand rows, then K3 / K4 / K3' unchanged."""
        clamp_mode, nerf_noise = kwargs["clamp_mode"], kwargs["nerf_noise"]
        white_back, last_back = kwargs.get("white_back", False), kwargs.get("last_back", False)
        draws = kwargs.get("draws") or {}
        B, R, dev = cam2worlds.shape[0], int(img_size) ** 2, cam2worlds.device

        def draw(name, fn, shape):
            t = draws.get(name)
            return fn(shape, device=dev) if t is None else t.to(dev)

        out: Dict[str, torch.Tensor] = {}
        u_jitter = draw("u_jitter", torch.rand, (B, R, S, 1))
        t_c, pts_c = ops.raymarch_points_coarse(cam2worlds, rays_d_cam, t_lin, u_jitter, img_size, img_size)
        coarse = net.mlp(net.point_features(pts_c.view(B, R * S, 3)), freq, phase)
        if hierarchical_sample:
            noise_c = draw("noise_coarse", torch.randn, (B, R, S, 1))
            _, _, w_c = ops.composite_fwd(coarse.view(B, R, S, 4), t_c, noise_c, nerf_noise, clamp_mode)
            u_re = draw("u_resample", torch.rand, (B * R, S))
            t_f = ops.resample_from_coarse(t_c, w_c, u_re)
            pts_f = ops.raymarch_points_fine(cam2worlds, rays_d_cam, t_f, img_size, img_size)
            fine = net.mlp(net.point_features(pts_f.view(B, R * S, 3)), freq, phase)
            noise_f = draw("noise_final", torch.randn, (B, R, 2 * S, 1))
            res = ops.merge_composite(fine, coarse, t_f, t_c, noise_f, rays_d_cam, B, img_size, img_size, nerf_noise, clamp_mode, white_back,
                                      last_back, taps=taps)
            if taps:
                out.update(weights_coarse=w_c, t_fine=t_f.view(B, R, S), points_fine=pts_f, rgb_sigma_fine=fine.view(B, R, S, 4))
        else:
            noise_f = draw("noise_final" if "noise_final" in draws else "noise_coarse", torch.randn, (B, R, S, 1))
            res = ops.merge_composite(None, coarse, None, t_c, noise_f, rays_d_cam, B, img_size, img_size, nerf_noise, clamp_mode, white_back,
                                      last_back, taps=taps)
        out["pixels"], out["depth"] = res[0], res[1]
        if taps:
            out.update(points_coarse=pts_c, t_coarse=t_c, rgb_sigma_coarse=coarse.view(B, R, S, 4), rgb=res[2]["rgb"], dist=res[2]["dist"],
                       merge_order=res[2]["order"])
        return out

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def staged_forward(self, z, cam2worlds, img_size, fov, ray_start, ray_end, num_steps, hierarchical_sample,
                       max_batch_size: int = 8, **kwargs):
        """Chunked no-grad inference: ``cam2worlds`` [P,4,4] poses are rendered ``max_batch_size`` at a
        time.  ``z`` may hold one object ([1,C,D,H,W] / [1,z_dim], broadcast to every pose: the video
        loop of inference.py:441-486) or one per pose.  ``fov`` may be a float or a length-P sequence
        (per-frame fov sweep, inference.py:459).  ``nerf_noise`` is forced to 0."""
        volume, global_feature = self.siren.split_z(z)
        P = cam2worlds.shape[0]
        if getattr(self.siren, "library_mlp", False):
            return self._staged_forward_library(z, cam2worlds, img_size, fov, ray_start, ray_end, num_steps, hierarchical_sample, max_batch_size, kwargs)
        n_obj = global_feature.shape[0] if self.siren.latent else volume.shape[0]
        shared = n_obj == 1 and P > 1
        kwargs = dict(kwargs)
        kwargs["nerf_noise"] = 0
        kwargs.setdefault("clamp_mode", "relu")
        fovs = [float(fov)] * P if not hasattr(fov, "__len__") else [float(f) for f in fov]
        vol_cl = None if self.siren.latent else ops.volume_to_channels_last(volume)
        film = self.siren.film_parameters(global_feature, n_obj, cam2worlds.device)
        pixels = torch.empty((P, 3, img_size, img_size), dtype=torch.float32, device=cam2worlds.device)
        depth = torch.empty((P, img_size, img_size), dtype=torch.float32, device=cam2worlds.device)
        start = 0
        while start < P:
            stop = min(start + max_batch_size, P)
            while stop > start + 1 and fovs[stop - 1] != fovs[start]:      # one fov per launch
                stop -= 1
            n = stop - start
            if shared:
                v = vol_cl                                                 # K1 reads item 0's volume for every pose (stride 0)
                f = tuple(t.expand(n, -1).contiguous() for t in film)
            else:
                v, f = (vol_cl[start:stop] if vol_cl is not None else None), tuple(t[start:stop] for t in film)
            o = self._render(None, None, cam2worlds[start:stop], img_size, fovs[start], ray_start, ray_end, num_steps,
                             hierarchical_sample, kwargs, vol_cl=v, film=f)
            pixels[start:stop], depth[start:stop] = o["pixels"], o["depth"]
            start = stop
        return pixels, depth

    def _staged_forward_library(self, z, cam2worlds, img_size, fov, ray_start, ray_end, num_steps, hierarchical_sample, max_batch_size, kwargs):
        """``staged_forward`` for the library-MLP decoders: pose chunks through ``autograd.render_library``; one object is expanded
        to the chunk's poses."""
        from .autograd import render_library
        P = cam2worlds.shape[0]
        kwargs = dict(kwargs)
        kwargs["nerf_noise"] = 0
        kwargs.setdefault("clamp_mode", "relu")
        fovs = [float(fov)] * P if not hasattr(fov, "__len__") else [float(f) for f in fov]

        def take(t, start, stop):
            if isinstance(t, (list, tuple)):
                return type(t)(take(x, start, stop) for x in t)
            if t.shape[0] == 1 and P > 1:
                return t.expand(stop - start, *t.shape[1:]).contiguous()
            return t[start:stop]

        pixels = torch.empty((P, 3, img_size, img_size), dtype=torch.float32, device=cam2worlds.device)
        depth = torch.empty((P, img_size, img_size), dtype=torch.float32, device=cam2worlds.device)
        start = 0
        while start < P:
            stop = min(start + max_batch_size, P)
            while stop > start + 1 and fovs[stop - 1] != fovs[start]:
                stop -= 1
            px, dp = render_library(self, take(z, start, stop), cam2worlds[start:stop], img_size, fovs[start], ray_start, ray_end, num_steps,
                                    hierarchical_sample, kwargs)
            pixels[start:stop], depth[start:stop] = px, dp
            start = stop
        return pixels, depth

    def generate_avg_frequencies(self):
        """generators/generators.py:189-197 expects a pi-GAN style mapping network returning a
        (frequencies, phase_shifts) pair; for the FG family it reduces to the FiLM parameters of random
        global features."""
        z = torch.randn((10000, self.z_dim), device=self.siren.mapping_network.weight.device)
        with torch.no_grad():
            frequencies, phase_shifts = self.siren.film_parameters(z)
        self.avg_frequencies = frequencies.mean(0, keepdim=True)
        self.avg_phase_shifts = phase_shifts.mean(0, keepdim=True)
        return self.avg_frequencies, self.avg_phase_shifts
