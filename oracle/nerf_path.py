"""Torch-CPU fp32 restatement of the reference rendering hot path (TEST INFRASTRUCTURE ONLY).

Every public function cites the reference lines it restates (paths relative to the reference
checkout, e.g. ``generators/volumetric_rendering.py:73-100``).  The functions take every random
draw as an explicit argument so that the CUDA path and the oracle consume identical numbers.

Parity status: PINNED against the unmodified reference executed in the build container
(``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``, checked by
``tests/test_oracle_golden.py``).  The reference itself has no tests or golden vectors.

Two deliberate, documented deviations from "whatever torch happens to do on this host":

* ``resample_pdf`` computes ``sum(weights)`` by accumulating in float64 and rounding once to
  float32.  The reference calls ``torch.sum`` whose fp32 association order differs between
  AVX2 / AVX512 / CUDA builds, i.e. the reference does not define the last bit.  The order
  independent value is the only one a second implementation can reproduce bit for bit.  The CDF
  is accumulated in float64 and rounded per element, which is exactly what ``torch.cumsum`` does
  for float32 on CPU (``at::acc_type<float, /*is_cuda=*/false>`` is double).
* ``merge_by_depth`` uses a stable sort (ties keep concatenation order: fine before coarse);
  ``torch.sort`` without ``stable=True`` leaves tie order unspecified.
"""

from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "VOXEL_LENGTH",
    "SIREN_SPECS",
    "layer_keys",
    "resolve_siren_type",
    "init_generator_state",
    "dense_head_state",
    "DENSE_HEAD_GAINS",
    "camera_rays",
    "jitter_samples",
    "camera_to_world",
    "random_camera_origins",
    "look_at_cam2world",
    "trilinear_lookup",
    "trilinear_manual",
    "film_parameters",
    "film_siren_mlp",
    "siren_forward",
    "composite",
    "resample_pdf",
    "coarse_to_fine_t",
    "fine_points",
    "merge_by_depth",
    "draw_randoms",
    "render",
    "psnr",
    "far_plane_sigma",
    "render_with_grad",
    "dense_grid_samples",
    "video_camera_origins",
]

# generators/siren.py:555 (and every other feature-volume variant): the voxel grid spans the
# cube [-0.6, 0.6]^3, so world coordinates are divided by voxel_length/2 before grid_sample.
VOXEL_LENGTH = 1.2

# generators/siren.py:491-538 (TALLSIREN_FG), :583-626 (SHORTSIREN_FG), :744-785
# (DOUBLESIREN_FG), :983-1023 (SingleSIREN_dg): number of FiLM layers, frequency_init divisor,
# whether the head applies _sigmoid_rgb (:579, :667, :826) or returns raw rgb (:1064).
SIREN_SPECS = {
    "TALLSIREN_FG": {"layers": 8, "freq_init": 25.0, "sigmoid_rgb": True},
    "SHORTSIREN_FG": {"layers": 4, "freq_init": 12.0, "sigmoid_rgb": True},
    "DOUBLESIREN_FG": {"layers": 2, "freq_init": 12.0, "sigmoid_rgb": True},
    "SingleSIREN_dg": {"layers": 1, "freq_init": 25.0, "sigmoid_rgb": False},
    # generators/siren.py:830-904: feature volume only, plain SirenLayer = sin(W x + b) (:180-199), no mapping network;
    # z is the feature volume itself (no global feature)
    "SHORTSIREN_F": {"layers": 4, "freq_init": 12.0, "sigmoid_rgb": True, "film": False},
    # generators/siren.py:333-408: SirenLayer, ResSirenBlock x2 (:218-230: y = sin(x + fc2(sin(fc1 x)))), SirenLayer, raw head;
    # feature volume only.  "layers" counts the linear layers in execution order; bit l of res_save: layer l's output is the
    # block input "x" kept for later, bit l of res_add: the kept x is added to layer l's pre-activation.
    "TALLSIREN_dRes": {"layers": 6, "freq_init": 25.0, "sigmoid_rgb": False, "film": False, "res_save": 0b000101, "res_add": 0b010100,
                       "keys": ["network.0.layer", "network.1.fc1", "network.1.fc2", "network.2.fc1", "network.2.fc2", "network.3.layer"]},
    # :411-488: the same with four residual blocks
    "TALLSIREN_dResLong": {"layers": 10, "freq_init": 25.0, "sigmoid_rgb": False, "film": False, "res_save": 0b0001010101, "res_add": 0b0101010100,
                           "keys": ["network.0.layer"] + [f"network.{b}.fc{i}" for b in (1, 2, 3, 4) for i in (1, 2)] + ["network.5.layer"]},
    # :906-979: one residual block, frequency_init(12), sigmoid on rgb
    "SHORTSIREN_FRes": {"layers": 4, "freq_init": 12.0, "sigmoid_rgb": True, "film": False, "res_save": 0b0001, "res_add": 0b0100,
                        "keys": ["network.0.layer", "network.1.fc1", "network.1.fc2", "network.2.layer"]},
    # generators/siren.py:1172-1224 (the default of configs/thousand/special.py:46): no feature volume -- the input of layer 0 is
    # the world position itself (input_dim = 3), z is a latent vector [B, z_dim] (PointNet encoder) mapped to the FiLM parameters
    # by CustomMappingNetwork (:55-78: three hidden Linear + LeakyReLU(0.2), kaiming_leaky_init, last weight x 0.25)
    "SHORTSIREN": {"layers": 4, "freq_init": 25.0, "sigmoid_rgb": True, "latent": True},
    # generators/siren.py:232-331: FiLM layers on the POSITION; frequencies / phases per point from PointFeaturesMappingNetwork
    # (:81-101) on the point's trilinear features; z = the feature volume alone; raw head (:326)
    "TALLSIREN": {"layers": 8, "freq_init": 25.0, "sigmoid_rgb": False, "pointwise": True},
    # :1068-1169: layer 0 reads cat([features, position]) (:1153); per-item FiLM from the global feature; raw head (:1164)
    "TALLSIREN_dgx": {"layers": 8, "freq_init": 25.0, "sigmoid_rgb": False, "xyz": True},
    # :671-741 + feature_pyramid_interpolation :1444-1473: layer 0 reads the concatenated trilinear features of every pyramid level
    "SHORTSIREN_FG_Pyrmd": {"layers": 4, "freq_init": 12.0, "sigmoid_rgb": True, "pyramid": True},
}


def layer_keys(siren_type: str):
    """State-dict prefixes (below ``siren.``) of the linear layers in execution order."""
    spec = SIREN_SPECS[resolve_siren_type(siren_type)]
    return spec.get("keys") or [f"network.{i}.layer" for i in range(spec["layers"])]

# configs/thousand/direct_volume/dg.py:8,51,55,59 spell the classes differently from
# generators/siren.py.
_SIREN_ALIASES = {
    "TALLSIREN_dg": "TALLSIREN_FG",
    "SHORTSIREN_dg": "SHORTSIREN_FG",
    "DoubleSIREN_dg": "DOUBLESIREN_FG",
    "DOUBLESIREN_dg": "DOUBLESIREN_FG",
}


def resolve_siren_type(name: str) -> str:
    name = _SIREN_ALIASES.get(name, name)
    if name not in SIREN_SPECS:
        raise AttributeError(f"module 'siren' has no attribute {name!r}")  # generators.py:15
    return name


def init_generator_state(
    siren_type: str, z_dim: int = 256, input_dim: int = 32, hidden_dim: int = 256, seed: int = 0
) -> Dict[str, torch.Tensor]:
    """Random-init parameters with the reference distributions and state_dict key names.

    generators/siren.py:134-143 (frequency_init), :40-44 (first_layer_film_sine_init), nn.Linear
    default init for biases and the mapping network; key names as produced by
    ``ImplicitGenerator3d.state_dict()`` (``siren.network.{i}.layer.weight`` ...).
    """
    spec = SIREN_SPECS[resolve_siren_type(siren_type)]
    g = torch.Generator().manual_seed(seed)

    def uniform(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    state: Dict[str, torch.Tensor] = {}
    for i, key in enumerate(layer_keys(siren_type)):
        fan_in = input_dim if i == 0 else hidden_dim
        w_bound = 1.0 / fan_in if i == 0 else math.sqrt(6.0 / fan_in) / spec["freq_init"]
        state[f"siren.{key}.weight"] = uniform((hidden_dim, fan_in), w_bound)
        state[f"siren.{key}.bias"] = uniform((hidden_dim,), 1.0 / math.sqrt(fan_in))
    state["siren.final_layer.weight"] = uniform((4, hidden_dim), math.sqrt(6.0 / hidden_dim) / spec["freq_init"])
    state["siren.final_layer.bias"] = uniform((4,), 1.0 / math.sqrt(hidden_dim))
    if not spec.get("film", True):
        return state
    if spec.get("pointwise"):
        # PointFeaturesMappingNetwork(z_dim, 256, L * hidden * 2): kaiming_normal_(a=0.2, fan_in), default biases, last weight x 0.25
        dims = [z_dim, 256, spec["layers"] * hidden_dim * 2]
        for i in range(2):
            std = math.sqrt(2.0 / (1 + 0.2 ** 2)) / math.sqrt(dims[i])
            w = torch.randn((dims[i + 1], dims[i]), generator=g) * std
            state[f"siren.mapping_network.network.{2 * i}.weight"] = w * (0.25 if i == 1 else 1.0)
            state[f"siren.mapping_network.network.{2 * i}.bias"] = uniform((dims[i + 1],), 1.0 / math.sqrt(dims[i]))
        return state
    if spec.get("latent"):
        # CustomMappingNetwork(z_dim, 256, L * hidden * 2): kaiming_normal_(a=0.2, fan_in) weights, default Linear biases, last weight x 0.25
        dims = [z_dim, 256, 256, 256, spec["layers"] * hidden_dim * 2]
        for i in range(4):
            std = math.sqrt(2.0 / (1 + 0.2 ** 2)) / math.sqrt(dims[i])
            w = torch.randn((dims[i + 1], dims[i]), generator=g) * std
            state[f"siren.mapping_network.network.{2 * i}.weight"] = w * (0.25 if i == 3 else 1.0)
            state[f"siren.mapping_network.network.{2 * i}.bias"] = uniform((dims[i + 1],), 1.0 / math.sqrt(dims[i]))
        return state
    n_map = spec["layers"] * hidden_dim * 2
    state["siren.mapping_network.weight"] = uniform((n_map, z_dim), 1.0 / math.sqrt(z_dim))
    state["siren.mapping_network.bias"] = uniform((n_map,), 1.0 / math.sqrt(z_dim))
    return state


# Random-init heads give sigma ~ 1e-2 (alpha ~ 1e-3 per sample, no occlusion, every relu-mode pixel decided by the far-plane
# sample).  The parity fixtures with REAL density scale the head rows of the SAME random-init network: (sigma gain, rgb gain),
# chosen per class so that alpha spans 0..~0.9 per sample, rays saturate before the far plane (mean far-plane weight < 1e-2)
# and the rendered image has texture.  The reference is run on these weights unchanged (tests/golden/make_golden.py).
# FiLM classes only: the unmodulated ones (SHORTSIREN_F, *_dRes*) are constant in space to ~1e-5 at random init whatever the
# head gain, so their image-level fixtures stay the near-empty ones (whole-image PSNR asserted in fp32 only).
# The sigma bias of +6 (TALLSIREN_FG) is a thin fog: no ray reaches the far plane with weight left (mean far-plane weight
# < 1e-3), so the far-plane step function (far_plane_sigma below) cannot decide a pixel; ~20 % of the samples stay empty.
DENSE_HEAD_GAINS = {           # class -> (sigma gain, rgb gain, first-layer gain, sigma bias)
    "TALLSIREN_FG": (300.0, 10.0, 1.0, 6.0),
    "SHORTSIREN_FG": (300.0, 3.0, 1.0, 0.0),
    "DOUBLESIREN_FG": (300.0, 3.0, 1.0, 0.0),
    "SingleSIREN_dg": (1000.0, 3.0, 1.0, 0.0),
}


def dense_head_state(state: Dict[str, torch.Tensor], sigma_gain: float, rgb_gain: float = 1.0, first_layer_gain: float = 1.0,
                     sigma_bias: float = 0.0) -> Dict[str, torch.Tensor]:
    """Copy of ``state`` whose head (``final_layer`` = nn.Linear(hidden, 4), siren.py:533) has its sigma row scaled by
    ``sigma_gain`` (bias set to ``sigma_bias``) and its three colour rows by ``rgb_gain``, and whose first linear layer is
    scaled by ``first_layer_gain``: a network with occlusion and texture."""
    st = {k: v.clone() for k, v in state.items()}
    st["siren.final_layer.weight"][3] *= sigma_gain
    st["siren.final_layer.bias"][3] = sigma_bias
    st["siren.final_layer.weight"][:3] *= rgb_gain
    if first_layer_gain != 1.0:
        first = "siren.network.0.layer.weight"
        st[first] *= first_layer_gain
    return st


# --------------------------------------------------------------------------------------------
# a1: camera-space rays and coarse sample distances
# --------------------------------------------------------------------------------------------
def camera_rays(
    batch: int, num_steps: int, img_size: int, fov: float, ray_start: float, ray_end: float
) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """generators/volumetric_rendering.py:73-100 (get_initial_rays_trig).

    Ray p = row * W + col looks through x = lin_W[col], y = lin_H[row] (no flip), z = 1/tan(fov/2)
    with tan evaluated in float64 (numpy) and divided into an fp32 ones tensor; directions are
    normalised; sample distances are ``linspace(ray_start, ray_end, num_steps)`` along the
    *normalised* ray.  Returns points[B,R,S,3], t[B,R,S,1], d_cam[B,R,3].
    """
    W = H = img_size
    lin_w = torch.linspace(-1, 1, W)
    lin_h = torch.linspace(-1, 1, H)
    x = lin_w.repeat(H)                      # x[row*W+col] = lin_w[col]
    y = lin_h.repeat_interleave(W)           # y[row*W+col] = lin_h[row]
    z = torch.ones_like(x) / np.tan((2 * math.pi * fov / 360) / 2)
    d = torch.stack([x, y, z], -1)
    d = d / torch.norm(d, dim=-1, keepdim=True)          # generators/math_utils_torch.py:16-20
    t = torch.linspace(ray_start, ray_end, num_steps).reshape(1, num_steps, 1).repeat(W * H, 1, 1)
    pts = d.unsqueeze(1).repeat(1, num_steps, 1) * t
    pts = pts.unsqueeze(0).repeat(batch, 1, 1, 1)
    t = t.unsqueeze(0).repeat(batch, 1, 1, 1)
    d = d.unsqueeze(0).repeat(batch, 1, 1)
    return pts, t, d


# --------------------------------------------------------------------------------------------
# a2: stratified jitter
# --------------------------------------------------------------------------------------------
def jitter_samples(pts, t, d_cam, u):
    """generators/volumetric_rendering.py:103-110 (perturb_points); ``u`` replaces torch.rand."""
    spacing = t[:, :, 1:2, :] - t[:, :, 0:1, :]
    offset = (u - 0.5) * spacing
    return pts + offset * d_cam.unsqueeze(2), t + offset


# --------------------------------------------------------------------------------------------
# a3: camera -> world
# --------------------------------------------------------------------------------------------
def camera_to_world(pts_cam, d_cam, cam2world):
    """generators/volumetric_rendering.py:161-192 (transform_sampled_points after the jitter).

    Homogeneous points through ``bmm(cam2world, P^T)``; directions through the 3x3 block; origin =
    cam2world applied to (0,0,0,1).  Returns pts_world[B,R,S,3], d_world[B,R,3], o_world[B,R,3].
    """
    B, R, S, _ = pts_cam.shape
    homo = torch.ones((B, R, S, 4), device=pts_cam.device)
    homo[..., :3] = pts_cam
    pts_w = torch.bmm(cam2world, homo.reshape(B, -1, 4).permute(0, 2, 1)).permute(0, 2, 1).reshape(B, R, S, 4)
    d_w = torch.bmm(cam2world[..., :3, :3], d_cam.reshape(B, -1, 3).permute(0, 2, 1)).permute(0, 2, 1).reshape(B, R, 3)
    o_h = torch.zeros((B, 4, R), device=pts_cam.device)
    o_h[:, 3, :] = 1
    o_w = torch.bmm(cam2world, o_h).permute(0, 2, 1).reshape(B, R, 4)[..., :3]
    return pts_w[..., :3], d_w, o_w


def random_camera_origins(n: int, r_start: float, r_end: float, up: str = "y", rng=None) -> torch.Tensor:
    """generators/volumetric_rendering.py:212-238 (sample_camera_positions), numpy RNG injectable."""
    assert up in ("y", "z")
    rng = np.random if rng is None else rng
    theta = np.clip(np.arccos(1 - rng.rand(n)), 1e-5, np.pi - 1e-5)
    phi = rng.rand(n) * np.pi * 2
    r = rng.rand(n) * (r_end - r_start) + r_start
    o = np.zeros((n, 3))
    o[:, 0] = r * np.sin(theta) * np.cos(phi)
    horiz, vert = r * np.sin(theta) * np.sin(phi), r * np.cos(theta)
    if up == "z":
        o[:, 1], o[:, 2] = horiz, vert
    else:
        o[:, 2], o[:, 1] = horiz, vert
    return torch.from_numpy(o).type(torch.float32)


def look_at_cam2world(origin: torch.Tensor, up: str = "y") -> torch.Tensor:
    """generators/volumetric_rendering.py:255-287 (create_cam2world_matrix)."""
    assert up in ("y", "z")

    def unit(v):
        return v / torch.norm(v, dim=-1, keepdim=True)

    fwd = unit(-origin)
    up_v = torch.tensor([0, 1, 0] if up == "y" else [0, 0, 1], dtype=torch.float).expand_as(fwd)
    left = unit(torch.cross(up_v, fwd, dim=-1))
    up_v = unit(torch.cross(fwd, left, dim=-1))
    n = origin.shape[0]
    rot = torch.eye(4).unsqueeze(0).repeat(n, 1, 1)
    rot[:, :3, :3] = torch.stack((-left, -up_v, fwd), dim=-1)
    trans = torch.eye(4).unsqueeze(0).repeat(n, 1, 1)
    trans[:, :3, 3] = origin
    return trans @ rot


# --------------------------------------------------------------------------------------------
# a4: trilinear feature lookup
# --------------------------------------------------------------------------------------------
def trilinear_lookup(volume, pts_world, img_size: int, num_steps: int):
    """generators/siren.py:555-571: grid_sample(bilinear, border, align_corners=False) -> [B,N,C]."""
    B, C = volume.shape[0], volume.shape[1]
    grid = (pts_world / (VOXEL_LENGTH / 2)).reshape(B, img_size, img_size, num_steps, 3)
    f = F.grid_sample(volume, grid, mode="bilinear", align_corners=False, padding_mode="border")
    return f.reshape(B, C, img_size**2 * num_steps).permute(0, 2, 1)


def trilinear_manual(volume: np.ndarray, pts_world: np.ndarray):
    """Explicit formula behind ``trilinear_lookup`` (ATen grid_sampler_3d, border padding).

    volume [C,D,H,W] fp32, pts_world [N,3] fp32.  Returns (features [N,C], corner index [N,3]
    int32 = floor of the clamped continuous index per axis, order x,y,z).  Used to pin the
    kernel's index arithmetic ("ray/sample indexing bit-exact").
    """
    f32 = np.float32
    C, D, H, W = volume.shape
    g = (pts_world.astype(f32) / f32(VOXEL_LENGTH / 2)).astype(f32)
    out_idx = np.zeros((pts_world.shape[0], 3), np.int32)
    lo, frac = [], []
    for axis, size in enumerate((W, H, D)):
        # unnormalise (align_corners=False): ((g + 1) * size - 1) / 2, then clip to [0, size-1]
        i = ((g[:, axis] + f32(1)) * f32(size) - f32(1)) / f32(2)
        i = np.minimum(np.maximum(i, f32(0)), f32(size - 1)).astype(f32)
        i0 = np.floor(i)
        out_idx[:, axis] = i0.astype(np.int32)
        lo.append(i0.astype(np.int64))
        frac.append((i - i0).astype(f32))
    x0, y0, z0 = lo
    fx, fy, fz = frac
    out = np.zeros((pts_world.shape[0], C), f32)
    for dz in (0, 1):            # ATen order: tnw, tne, tsw, tse, bnw, bne, bsw, bse
        for dy in (0, 1):
            for dx in (0, 1):
                wx = fx if dx else (f32(1) - fx)  # == (x0+1) - i in ATen, identical in fp32 here
                wy = fy if dy else (f32(1) - fy)
                wz = fz if dz else (f32(1) - fz)
                xi, yi, zi = x0 + dx, y0 + dy, z0 + dz
                inb = (xi <= W - 1) & (yi <= H - 1) & (zi <= D - 1)
                w = (wx * wy * wz).astype(f32) * inb
                v = volume[:, np.minimum(zi, D - 1), np.minimum(yi, H - 1), np.minimum(xi, W - 1)]
                out += (v * w[None, :]).T.astype(f32)
    return out, out_idx


# --------------------------------------------------------------------------------------------
# a5-a7: FiLM-SIREN MLP
# --------------------------------------------------------------------------------------------
def film_parameters(global_feature, map_weight, map_bias):
    """generators/siren.py:550-553: freq = first half * 15 + 30, phase = second half."""
    fo = F.linear(global_feature, map_weight, map_bias)
    half = fo.shape[-1] // 2
    return fo[..., :half] * 15 + 30, fo[..., half:]


def custom_mapping_network(z, state):
    """generators/siren.py:55-78 (CustomMappingNetwork.forward) + :1212-1215: freq = first half * 15 + 30, phase = second half."""
    x = z
    for i in range(4):
        x = F.linear(x, state[f"siren.mapping_network.network.{2 * i}.weight"], state[f"siren.mapping_network.network.{2 * i}.bias"])
        if i < 3:
            x = F.leaky_relu(x, 0.2)
    half = x.shape[-1] // 2
    return x[..., :half] * 15 + 30, x[..., half:]


def film_siren_mlp(feat, layer_weights, layer_biases, freq, phase, final_w, final_b, sigmoid_rgb=True, res_save=0, res_add=0):
    """generators/siren.py:573-579 + FiLMLayer.forward :153-160 + _sigmoid_rgb :1227-1234.

    feat [B,N,K0]; freq/phase [B, L*H]; returns rgb_sigma [B,N,4] (rgb through sigmoid iff
    ``sigmoid_rgb``, sigma raw).  Dropout has p=0 in every shipped config and is omitted.
    """
    x = feat
    H = layer_weights[0].shape[0]
    kept = None
    for i, (w, b) in enumerate(zip(layer_weights, layer_biases)):
        x = F.linear(x, w, b)
        fr = freq[:, i * H:(i + 1) * H].unsqueeze(1).expand_as(x)
        ph = phase[:, i * H:(i + 1) * H].unsqueeze(1).expand_as(x)
        u = fr * x + ph
        if (res_add >> i) & 1:
            u = kept + u                                 # ResSirenBlock.forward, siren.py:228: sin(x + net)
        x = torch.sin(u)
        if (res_save >> i) & 1:
            kept = x
    out = F.linear(x, final_w, final_b)
    if sigmoid_rgb:
        out = torch.cat([torch.sigmoid(out[..., :3]), out[..., -1:]], dim=-1)
    return out


def _split_state(state, siren_type):
    spec = SIREN_SPECS[resolve_siren_type(siren_type)]
    ws = [state[f"siren.{k}.weight"] for k in layer_keys(siren_type)]
    bs = [state[f"siren.{k}.bias"] for k in layer_keys(siren_type)]
    return spec, ws, bs


def siren_forward(state, siren_type, pts_world, z, img_size, num_steps):
    """``SIREN.forward(points, z, img_size, num_steps)`` for the FG family (siren.py:540-580)."""
    spec, ws, bs = _split_state(state, siren_type)
    fw_, fb_ = state["siren.final_layer.weight"], state["siren.final_layer.bias"]
    if spec.get("pointwise"):                        # siren.py:292-326
        feat = trilinear_lookup(z, pts_world, img_size, num_steps)
        h = F.leaky_relu(F.linear(feat, state["siren.mapping_network.network.0.weight"], state["siren.mapping_network.network.0.bias"]), 0.2)
        fo = F.linear(h, state["siren.mapping_network.network.2.weight"], state["siren.mapping_network.network.2.bias"])
        half = fo.shape[-1] // 2
        freq, phase = fo[..., :half] * 15 + 30, fo[..., half:]
        x, H = pts_world, ws[0].shape[0]
        for i, (w, b) in enumerate(zip(ws, bs)):      # PointwiseFiLMLayer.forward :170-177
            x = torch.sin(freq[..., i * H:(i + 1) * H] * F.linear(x, w, b) + phase[..., i * H:(i + 1) * H])
        return F.linear(x, fw_, fb_)
    if spec.get("xyz") or spec.get("pyramid"):
        volumes, global_feature = z
        freq, phase = film_parameters(global_feature, state["siren.mapping_network.weight"], state["siren.mapping_network.bias"])
        if spec.get("xyz"):                          # siren.py:1136-1153
            x0 = torch.cat([trilinear_lookup(volumes, pts_world, img_size, num_steps), pts_world], dim=-1)
        else:                                        # feature_pyramid_interpolation, siren.py:1444-1473
            levels = volumes if isinstance(volumes, (list, tuple)) else [volumes]
            x0 = torch.cat([trilinear_lookup(v, pts_world, img_size, num_steps) for v in levels], dim=2)
        return film_siren_mlp(x0, ws, bs, freq, phase, fw_, fb_, spec["sigmoid_rgb"])
    if spec.get("latent"):                           # siren.py:1206-1224: x = input points, z = latent vector
        freq, phase = custom_mapping_network(z, state)
        return film_siren_mlp(pts_world, ws, bs, freq, phase, state["siren.final_layer.weight"], state["siren.final_layer.bias"], spec["sigmoid_rgb"])
    if spec.get("film", True):
        volume, global_feature = z
        freq, phase = film_parameters(global_feature, state["siren.mapping_network.weight"], state["siren.mapping_network.bias"])
    else:
        volume = z                                   # siren.py:867: forward(points, feature_volume, ...); sin(1 * x + 0) == sin(x)
        n = spec["layers"] * ws[0].shape[0]
        freq, phase = torch.ones((volume.shape[0], n), device=volume.device), torch.zeros((volume.shape[0], n), device=volume.device)
    feat = trilinear_lookup(volume, pts_world, img_size, num_steps)
    return film_siren_mlp(feat, ws, bs, freq, phase, state["siren.final_layer.weight"],
                          state["siren.final_layer.bias"], spec["sigmoid_rgb"], spec.get("res_save", 0), spec.get("res_add", 0))


# --------------------------------------------------------------------------------------------
# a8: alpha compositing
# --------------------------------------------------------------------------------------------
def composite(rgb_sigma, t, noise, noise_std, clamp_mode, white_back=False, last_back=False):
    """generators/volumetric_rendering.py:18-70 (fancy_integration); ``noise`` replaces randn.

    rgb_sigma [B,R,S,4], t [B,R,S,1], noise [B,R,S,1].  Returns rgb[B,R,3], dist[B,R,1],
    weights[B,R,S,1].  (fill_mode is unused by every caller and not restated.)
    """
    rgb, sigma = rgb_sigma[..., :3], rgb_sigma[..., 3:]
    delta = t[:, :, 1:] - t[:, :, :-1]
    delta = torch.cat([delta, 1e10 * torch.ones_like(delta[:, :, :1])], -2)
    s = sigma + noise * noise_std
    if clamp_mode == "softplus":
        s = F.softplus(s)
    elif clamp_mode == "relu":
        s = F.relu(s)
    else:
        raise TypeError("Need to choose clamp mode")  # reference: `raise "<str>"` -> TypeError
    alpha = 1 - torch.exp(-delta * s)
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :, :1]), 1 - alpha + 1e-10], -2), -2)[:, :, :-1]
    w = alpha * trans
    w_sum = w.sum(2)
    if last_back:
        w[:, :, -1] += 1 - w_sum
    rgb_out = torch.sum(w * rgb, -2)
    dist = torch.sum(w * t, -2)
    if white_back:
        rgb_out = rgb_out + 1 - w_sum
    return rgb_out, dist, w


# --------------------------------------------------------------------------------------------
# a9: inverse-CDF importance resampling
# --------------------------------------------------------------------------------------------
def resample_pdf(bins, weights, u, eps: float = 1e-5):
    """generators/volumetric_rendering.py:297-342 (sample_pdf, det=False); ``u`` replaces rand.

    bins [N,M+1], weights [N,M], u [N,K].  Returns (samples [N,K] fp32, inds, below, above int64).
    See the module docstring for the accumulation-order convention.
    """
    M = weights.shape[1]
    w = weights + eps
    total = w.double().sum(-1, keepdim=True).float()
    pdf = w / total
    cdf = torch.cumsum(pdf.double(), -1).float()
    cdf = torch.cat([torch.zeros_like(cdf[:, :1]), cdf], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u)                     # right=False: first i with cdf[i] >= u
    below = torch.clamp_min(inds - 1, 0)
    above = torch.clamp_max(inds, M)
    cdf_lo, cdf_hi = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    bin_lo, bin_hi = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = cdf_hi - cdf_lo
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    samples = bin_lo + (u - cdf_lo) / denom * (bin_hi - bin_lo)
    return samples, inds, below, above


def coarse_to_fine_t(weights, t, u, num_steps):
    """generators/generators.py:123-136: the call site of sample_pdf (two separate +1e-5)."""
    N = weights.shape[0] * weights.shape[1]
    w = weights.reshape(N, num_steps) + 1e-5
    tt = t.reshape(N, num_steps)
    mid = 0.5 * (tt[:, :-1] + tt[:, 1:])
    return resample_pdf(mid, w[:, 1:-1], u)


# --------------------------------------------------------------------------------------------
# a10, a11
# --------------------------------------------------------------------------------------------
def fine_points(o_world, d_world, t_fine):
    """generators/generators.py:138-145: p = o + d * t (world space)."""
    return o_world.unsqueeze(2).contiguous() + d_world.unsqueeze(2).contiguous() * t_fine.expand(-1, -1, -1, 3).contiguous()


def merge_by_depth(fine_out, coarse_out, t_fine, t_coarse):
    """generators/generators.py:163-167: concatenate (fine first), sort by t, gather."""
    all_out = torch.cat([fine_out, coarse_out], dim=-2)
    all_t = torch.cat([t_fine, t_coarse], dim=-2)
    _, order = torch.sort(all_t, dim=-2, stable=True)
    return torch.gather(all_out, -2, order.expand(-1, -1, -1, 4)), torch.gather(all_t, -2, order), order


# --------------------------------------------------------------------------------------------
# whole forward
# --------------------------------------------------------------------------------------------
def draw_randoms(batch, img_size, num_steps, hierarchical=True, generator: Optional[torch.Generator] = None):
    """The four draws of one forward, in the reference's order and shapes:
    rand[B,R,S,1], randn[B,R,S,1], rand[B*R,S], randn[B,R,2S,1]."""
    R = img_size * img_size
    d = {
        "u_jitter": torch.rand((batch, R, num_steps, 1), generator=generator),
        "noise_coarse": torch.randn((batch, R, num_steps, 1), generator=generator),
    }
    if hierarchical:
        d["u_resample"] = torch.rand((batch * R, num_steps), generator=generator)
        d["noise_final"] = torch.randn((batch, R, 2 * num_steps, 1), generator=generator)
    return d


def render(state, siren_type, z, cam2worlds, draws, *, img_size, fov, ray_start, ray_end, num_steps,
           hierarchical_sample, clamp_mode, nerf_noise, white_back=False, last_back=False,
           taps: bool = True, **_ignored) -> Dict[str, torch.Tensor]:
    """generators/generators.py:33-187 (ImplicitGenerator3d.forward) with replayed draws.

    Returns a dict with ``pixels`` [B,3,H,W], ``depth`` [B,H,W] and (``taps=True``) every
    intermediate the parity tests compare against.  Runs without autograd; ``render_with_grad``
    is the same code with the reference's grad / no-grad structure.
    """
    with torch.no_grad():
        return _render(state, siren_type, z, cam2worlds, draws, img_size=img_size, fov=fov, ray_start=ray_start, ray_end=ray_end,
                       num_steps=num_steps, hierarchical_sample=hierarchical_sample, clamp_mode=clamp_mode, nerf_noise=nerf_noise,
                       white_back=white_back, last_back=last_back, taps=taps)


def render_with_grad(state, siren_type, z, cam2worlds, draws, **meta):
    """As ``render`` but differentiable w.r.t. ``state`` tensors, the volume and the global feature, with the
    reference's structure: rays and resampling under torch.no_grad() (generators.py:57, :111), both SIREN
    passes and the final composite with grad (:102-107, :155-160, :172-180)."""
    meta = {k: v for k, v in meta.items() if k in ("img_size", "fov", "ray_start", "ray_end", "num_steps", "hierarchical_sample",
                                                    "clamp_mode", "nerf_noise", "white_back", "last_back")}
    return _render(state, siren_type, z, cam2worlds, draws, taps=False, **meta)


def _render(state, siren_type, z, cam2worlds, draws, *, img_size, fov, ray_start, ray_end, num_steps,
            hierarchical_sample, clamp_mode, nerf_noise, white_back=False, last_back=False, taps=True):
    B = cam2worlds.shape[0]
    R, S = img_size * img_size, num_steps
    out: Dict[str, torch.Tensor] = {}
    with torch.no_grad():
        # the reference builds its rays on the generator's device (generators.py:57-66); here: where the cameras live
        pts_cam, t, d_cam = (x.to(cam2worlds.device) for x in camera_rays(B, S, img_size, fov, ray_start, ray_end))
        pts_cam, t = jitter_samples(pts_cam, t, d_cam, draws["u_jitter"])
        pts_w, d_w, o_w = camera_to_world(pts_cam, d_cam, cam2worlds)
    coarse = siren_forward(state, siren_type, pts_w.reshape(B, R * S, 3), z, img_size, S).reshape(B, R, S, 4)
    if taps:
        out.update(points_coarse=pts_w, t_coarse=t, dirs_world=d_w, origins_world=o_w, rgb_sigma_coarse=coarse)
    if hierarchical_sample:
        with torch.no_grad():
            _, _, w = composite(coarse, t, draws["noise_coarse"], nerf_noise, clamp_mode)
            t_fine, inds, below, above = coarse_to_fine_t(w, t, draws["u_resample"], S)
            t_fine = t_fine.reshape(B, R, S, 1)
            pts_f = fine_points(o_w, d_w, t_fine)
        fine = siren_forward(state, siren_type, pts_f.reshape(B, R * S, 3), z, img_size, S).reshape(B, R, -1, 4)
        all_out, all_t, order = merge_by_depth(fine, coarse, t_fine, t)
        final_noise = draws["noise_final"]
        if taps:
            out.update(weights_coarse=w, t_fine=t_fine, resample_inds=inds, points_fine=pts_f,
                       rgb_sigma_fine=fine, merge_order=order, t_all=all_t)
    else:
        all_out, all_t = coarse, t
        # generators.py:172-180: the final composite draws a second randn of the same shape
        final_noise = draws.get("noise_final", draws["noise_coarse"])
    rgb, dist, w_all = composite(all_out, all_t, final_noise, nerf_noise, clamp_mode,
                                 white_back=white_back, last_back=last_back)
    out["pixels"] = rgb.reshape(B, img_size, img_size, 3).permute(0, 3, 1, 2).contiguous() * 2 - 1
    out["depth"] = (d_cam[..., -1:] * dist).reshape(B, img_size, img_size).contiguous()  # vr.py:345-356
    if taps:
        out.update(rgb=rgb, dist=dist, weights_final=w_all, rgb_sigma_all=all_out, noise_final=final_noise)
    return out


def far_plane_sigma(out: Dict[str, torch.Tensor], nerf_noise: float) -> torch.Tensor:
    """Density (+ noise) of the farthest composited sample of every ray, [B, R], from ``render(taps=True)``.

    generators/volumetric_rendering.py:34-35 gives that sample delta = 1e10, so with clamp_mode "relu" its
    alpha is the STEP function [sigma > 0] (1 - exp(-1e10 * relu(sigma))): the reference image is
    discontinuous in that one value, and a pixel whose far-plane |sigma| is below the arithmetic tolerance
    of the MLP (bf16 here, fp16 under the reference's own autocast) is decided by the sign of a number that
    is zero within tolerance.  Image-level comparisons of a reduced-precision path therefore either use
    clamp_mode "softplus" (alpha_last == 1 always) or exclude those pixels and report their fraction.
    """
    return out["rgb_sigma_all"][:, :, -1, 3] + out["noise_final"][:, :, -1, 0] * nerf_noise


LIBRARY_CASES = {       # siren_type -> (z_dim, input_dim, channels of the volume(s), seed) of the fixtures tests/golden/fwd_<type>.npz
    "TALLSIREN": (32, 3, [32], 51),
    "TALLSIREN_dgx": (256, 35, [32], 52),
    "SHORTSIREN_FG_Pyrmd": (256, 224, [32, 64, 128], 53),
}


def library_case_inputs(siren_type: str, B: int = 2, V: int = 12):
    """Seeded (state, z, cam2world, generator) of the three decoders above: z = the volume alone (TALLSIREN), (volume, global)
    (TALLSIREN_dgx) or (list of pyramid volumes at V, V/2, V/4, global) (SHORTSIREN_FG_Pyrmd)."""
    z_dim, input_dim, plan, seed = LIBRARY_CASES[siren_type]
    g = torch.Generator().manual_seed(seed)
    state = init_generator_state(siren_type, z_dim=z_dim, input_dim=input_dim, hidden_dim=256, seed=seed)
    vols = [torch.randn((B, c, max(V >> i, 2), max(V >> i, 2), max(V >> i, 2)), generator=g) * 0.3 for i, c in enumerate(plan)]
    glob = torch.randn((B, 256), generator=g) * 0.05 + 0.19
    if siren_type == "TALLSIREN":
        z = vols[0]
    elif siren_type == "TALLSIREN_dgx":
        z = (vols[0], glob)
    else:
        z = (vols, glob)
    cam = look_at_cam2world(random_camera_origins(B, 0.7, 1.5, "y", np.random.RandomState(seed)), "y")
    return state, z, cam, g


def dense_grid_samples(N: int, voxel_origin=(0, 0, 0), cube_length: float = 2.0):
    """extract_shapes.py:15-37 (create_samples): [1, N^3, 3] sample positions, z fastest; the x / y indices come from a
    float division and are deliberately left non-integer, as in the reference."""
    corner = np.array(voxel_origin) - cube_length / 2
    voxel_size = cube_length / (N - 1)
    idx = torch.arange(0, N ** 3, 1, dtype=torch.int64)
    s = torch.zeros(N ** 3, 3)
    s[:, 2] = idx % N
    s[:, 1] = (idx.float() / N) % N
    s[:, 0] = ((idx.float() / N) / N) % N
    s[:, 0] = (s[:, 0] * voxel_size) + corner[2]
    s[:, 1] = (s[:, 1] * voxel_size) + corner[1]
    s[:, 2] = (s[:, 2] * voxel_size) + corner[0]
    return s.unsqueeze(0), corner, voxel_size


def video_camera_origins(num_frames: int, fps: int, r_start: float, r_end: float, up: str = "y"):
    """inference.py:442-470: camera origins [F,3] (float32 tensor) and the fov ramp [F] of ``render_video``."""
    theta0 = np.linspace(1e-5, np.pi / 2 - 1e-5, num_frames // 2)
    phi0 = np.linspace(0, np.pi * 2, num_frames // 2)
    theta1 = np.linspace(np.pi / 2 - 1e-5, 1e-5, num_frames // 4)
    phi11 = np.linspace(np.pi * 2, np.pi * 5 / 4, fps)
    phi12 = np.asarray([np.pi * 5 / 4] * (num_frames // 4 - fps))
    theta21 = np.linspace(1e-5, np.pi / 4 - 1e-5, fps)
    theta22 = np.asarray([np.pi / 4 - 1e-5] * (num_frames // 4 - fps))
    phi2 = np.linspace(np.pi * 5 / 4, 0, num_frames // 4)
    theta = np.concatenate([theta0, theta1, theta21, theta22], axis=0)
    phi = np.concatenate([phi0, phi11, phi12, phi2], axis=0)
    r = np.linspace(r_start, r_end, num_frames)
    fov = np.linspace(60, 30, num_frames)
    o = np.zeros((num_frames, 3))
    o[:, 0] = r * np.sin(theta) * np.cos(phi)
    if up == "z":
        o[:, 1] = r * np.sin(theta) * np.sin(phi)
        o[:, 2] = r * np.cos(theta)
    else:
        o[:, 2] = r * np.sin(theta) * np.sin(phi)
        o[:, 1] = r * np.cos(theta)
    return torch.from_numpy(o).type(torch.float32), fov


def psnr(a: torch.Tensor, b: torch.Tensor, data_range: float = 2.0) -> float:
    """metric_utils.py:245-256 style PSNR; images live in [-1,1] so the range is 2."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return float("inf") if mse == 0 else 10.0 * math.log10(data_range**2 / mse)
