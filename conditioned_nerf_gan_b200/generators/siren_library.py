"""The three SIREN decoders of the reference that are NOT on the fused tcgen05 kernels: ``TALLSIREN`` (per-point FiLM,
generators/siren.py:232-331, selected by configs/thousand/direct_volume/indirect.py:10), ``TALLSIREN_dgx`` (features concatenated
with the position, :1068-1169) and ``SHORTSIREN_FG_Pyrmd`` (multi-resolution feature pyramid, :671-741, :1444-1473).

They are provided so that every class a reference config or checkpoint can name is a working drop-in (same constructor, same
state-dict keys, same ``forward(points, z, img_size, num_steps)``), with this split:

  * the trilinear lookups, the ray generation, compositing, resampling, merge and their backward kernels are the library's
    (``cng_gather_points`` / ``cng_scatter_points`` through ``autograd._GatherPoints``, K1 in points-only mode, K3 / K4 / K3');
  * the MLP itself is **plain PyTorch** (``F.linear`` + ``torch.sin``: library GEMMs, fp32 or the caller's autocast dtype),
    evaluated in chunks of points -- and under ``torch.utils.checkpoint`` when gradients are needed -- so that the per-point
    tensors the reference materialises for a whole image ([B, N, 4096] FiLM parameters for ``TALLSIREN``: 16 KB per point) never
    exceed one chunk.

Why not the fused kernels: ``TALLSIREN`` needs three contractions per layer and point (W_l x, and the frequency / phase rows of
the per-point mapping network) and their elementwise combination, i.e. three TMEM accumulators per layer; ``_dgx`` has 35 inputs
and ``_Pyrmd`` 224 (the fused kernels' layer 0 is built for the 32-channel split hi/lo operand).  DESIGN.md section 6.
"""
from __future__ import annotations

import math
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F
import torch.utils.checkpoint

from .. import ops
from .siren import FiLMLayer, _uniform_

__all__ = ["PointFeaturesMappingNetwork", "PointwiseFiLMLayer", "TALLSIREN", "TALLSIREN_dgx", "SHORTSIREN_FG_Pyrmd"]

CHUNK_POINTS = 1 << 16          # points per evaluation chunk of the library MLP (TALLSIREN: 1 GB of fp32 FiLM parameters per chunk)

PointwiseFiLMLayer = FiLMLayer   # siren.py:163-177: the same single nn.Linear (+ parameter-free dropout); the FiLM arithmetic lives in _body


class PointFeaturesMappingNetwork(nn.Module):
    """generators/siren.py:81-101: per-point features [.., z_dim] -> (frequencies, phase shifts) [.., out/2] each."""

    def __init__(self, z_dim: int, map_hidden_dim: int, map_output_dim: int):
        super().__init__()
        self.network = nn.Sequential(nn.Linear(z_dim, map_hidden_dim), nn.LeakyReLU(0.2, inplace=True), nn.Linear(map_hidden_dim, map_output_dim))
        for m in self.network:
            if isinstance(m, nn.Linear):
                torch.nn.init.kaiming_normal_(m.weight, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
        with torch.no_grad():
            self.network[-1].weight *= 0.25

    def forward(self, z):
        fo = self.network(z)
        half = fo.shape[-1] // 2
        return fo[..., :half], fo[..., half:]


def _gather(volume: torch.Tensor, points: torch.Tensor) -> torch.Tensor:
    """Trilinear lookup [B,C,D,H,W] x [B,N,3] -> [B,N,C] on the library's kernels, differentiable w.r.t. the volume."""
    from .autograd import _GatherPoints, _ToChannelsLast
    if torch.is_grad_enabled() and volume.requires_grad:
        return _GatherPoints.apply(_ToChannelsLast.apply(volume.float()), points.detach())
    return ops.gather_points(ops.volume_to_channels_last(volume), points)


class _LibrarySiren(nn.Module):
    """Shared host logic: chunked (and, with grad, checkpointed) evaluation of ``_body`` over the points."""
    library_mlp = True
    latent = False
    film = True
    num_layers = 0
    freq_div = 25.0
    sigmoid_rgb = False
    res_add_mask = 0
    res_save_mask = 0
    precision = "fp32"              # informational: the MLP is a torch op chain (fp32, or the caller's autocast dtype)

    def _init_weights(self, first_in: int):
        for i, film in enumerate(self.network):
            fan_in = film.layer.weight.shape[-1]
            _uniform_(film.layer, 1.0 / fan_in if i == 0 else math.sqrt(6.0 / fan_in) / self.freq_div)
        _uniform_(self.final_layer, math.sqrt(6.0 / self.hidden_dim) / self.freq_div)

    def check_dropout(self) -> None:
        if self.training and any(getattr(m, "drop_out_prob", 0) > 0 for m in self.network):
            raise NotImplementedError("FiLM dropout > 0 in training mode is not built (call .eval(), or construct with drop_out=0)")

    def _layers(self, x, freq, phase):
        """x [P,K0]; freq / phase [P or 1, L*H] -> rgb_sigma [P,4] (FiLMLayer.forward :153-160 / PointwiseFiLMLayer :170-177)."""
        H = self.hidden_dim
        for i, film in enumerate(self.network):
            x = torch.sin(freq[..., i * H:(i + 1) * H] * film.layer(x) + phase[..., i * H:(i + 1) * H])
        out = self.final_layer(x)
        if self.sigmoid_rgb:
            out = torch.cat([torch.sigmoid(out[..., :3]), out[..., 3:]], dim=-1)
        return out

    def _chunked(self, fn, n_points: int, *per_point):
        """Evaluate ``fn(*slices)`` over chunks of the point axis (dim 1 of every tensor in ``per_point``)."""
        outs = []
        use_ckpt = torch.is_grad_enabled()
        for s in range(0, n_points, CHUNK_POINTS):
            sl = [t[:, s:s + CHUNK_POINTS] for t in per_point]
            outs.append(torch.utils.checkpoint.checkpoint(fn, *sl, use_reentrant=False) if use_ckpt else fn(*sl))
        return torch.cat(outs, dim=1) if len(outs) > 1 else outs[0]


class TALLSIREN(_LibrarySiren):
    """generators/siren.py:232-331: eight FiLM layers on the sample POSITION whose frequencies / phase shifts come, per point,
    from a mapping network on the point's trilinear features.  ``z`` is the feature volume alone; raw rgb head."""
    num_layers, freq_div, sigmoid_rgb = 8, 25.0, False

    def __init__(self, input_dim=3, z_dim=100, hidden_dim=256, output_dim=4, drop_out=0, device=None, **kwargs):
        super().__init__()
        self.device = device
        self.input_dim, self.z_dim, self.hidden_dim, self.output_dim = input_dim, z_dim, hidden_dim, output_dim
        self.network = nn.ModuleList([PointwiseFiLMLayer(input_dim if i == 0 else hidden_dim, hidden_dim, drop_out) for i in range(self.num_layers)])
        self.final_layer = nn.Linear(hidden_dim, 4)
        self.mapping_network = PointFeaturesMappingNetwork(z_dim, 256, self.num_layers * hidden_dim * 2)
        self._init_weights(input_dim)

    def split_z(self, z):
        if isinstance(z, (tuple, list)):
            raise ValueError("TALLSIREN takes the feature volume alone as z (generators/siren.py:292-299)")
        return z, None

    def forward(self, points: torch.Tensor, z, img_size: int = 0, num_steps: int = 0) -> torch.Tensor:
        self.check_dropout()
        volume, _ = self.split_z(z)
        feat = _gather(volume, points)                                   # [B, N, z_dim]

        def body(p, f):
            freq, phase = self.mapping_network(f)
            return self._layers(p, freq * 15 + 30, phase)

        return self._chunked(body, points.shape[1], points.float(), feat)


class TALLSIREN_dgx(_LibrarySiren):
    """generators/siren.py:1068-1169: the FG decoder whose layer 0 reads the trilinear features concatenated with the position
    (``input_dim`` = feature channels + 3); FiLM parameters per item from the global feature; raw rgb head."""
    num_layers, freq_div, sigmoid_rgb = 8, 25.0, False

    def __init__(self, input_dim=3, z_dim=100, hidden_dim=256, output_dim=4, drop_out=0, device=None, **kwargs):
        super().__init__()
        self.device = device
        self.input_dim, self.z_dim, self.hidden_dim, self.output_dim = input_dim, z_dim, hidden_dim, output_dim
        self.network = nn.ModuleList([FiLMLayer(input_dim if i == 0 else hidden_dim, hidden_dim, drop_out) for i in range(self.num_layers)])
        self.final_layer = nn.Linear(hidden_dim, 4)
        self.mapping_network = nn.Linear(z_dim, self.num_layers * hidden_dim * 2)
        self._init_weights(input_dim)

    def split_z(self, z):
        if not isinstance(z, (tuple, list)) or len(z) != 2:
            raise ValueError(f"{type(self).__name__} needs z = (feature volume(s), global_feature [B,z_dim])")
        return z[0], z[1]

    def _film(self, global_feature):
        fo = F.linear(global_feature.float(), self.mapping_network.weight.float(), self.mapping_network.bias.float())
        half = fo.shape[-1] // 2
        return (fo[..., :half] * 15 + 30).unsqueeze(1), fo[..., half:].unsqueeze(1)          # [B, 1, L*H]: broadcast over the points

    def _features(self, volume, points):
        return torch.cat([_gather(volume, points), points.float()], dim=-1)             # :1153

    def forward(self, points: torch.Tensor, z, img_size: int = 0, num_steps: int = 0) -> torch.Tensor:
        self.check_dropout()
        volume, global_feature = self.split_z(z)
        freq, phase = self._film(global_feature)
        x0 = self._features(volume, points)
        return self._chunked(lambda x: self._layers(x, freq, phase), points.shape[1], x0)


class SHORTSIREN_FG_Pyrmd(TALLSIREN_dgx):
    """generators/siren.py:671-741 with ``feature_pyramid_interpolation`` (:1444-1473): four FiLM layers on the concatenation of
    the trilinear features of every level of a feature pyramid (``PyramidUNet3D``; ``input_dim`` = the summed channels, 224 for
    32 + 64 + 128); frequency_init(12), sigmoid on rgb."""
    num_layers, freq_div, sigmoid_rgb = 4, 12.0, True

    def _features(self, pyramid, points):
        levels: List[torch.Tensor] = list(pyramid) if isinstance(pyramid, (list, tuple)) else [pyramid]
        return torch.cat([_gather(v, points) for v in levels], dim=2)
