"""Torch-tensor front end of the C ABI (include/cng_b200.h): device pointers, sizes and the
current CUDA stream go down through ctypes; tensors stay owned by PyTorch.

Every function here launches hand-written sm_100a kernels from libcng_b200.so.  There is no CPU
or eager-PyTorch fallback: a tensor that is not on a CUDA device raises, a missing library
raises at import of ``_lib.load()``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Tuple

import torch

from . import _lib

CLAMP_MODES = {"relu": _lib.CLAMP_RELU, "softplus": _lib.CLAMP_SOFTPLUS}
PRECISIONS = {"fp32": _lib.PREC_FP32, "bf16": _lib.PREC_BF16, "fp16": _lib.PREC_FP16}

# number of kernel launches issued through this module (bench.py reports it as gpu_launches)
launch_count = 0


def _count(n: int = 1) -> None:
    global launch_count
    launch_count += n


# Optional per-entry-point device timing (bench.py's roofline leg): when `kernel_events` is a dict,
# every C-ABI call is bracketed by CUDA events recorded on the launching stream.
kernel_events = None


class _timed:
    __slots__ = ("name", "start")

    def __init__(self, name: str):
        self.name = name
        self.start = None

    def __enter__(self):
        if kernel_events is not None:
            self.start = torch.cuda.Event(enable_timing=True)
            self.start.record()
        return self

    def __exit__(self, *exc):
        if self.start is not None:
            end = torch.cuda.Event(enable_timing=True)
            end.record()
            kernel_events.setdefault(self.name, []).append((self.start, end))
        return False


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tensor is on {t.device}; the rendering path has no CPU implementation")
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(t: torch.Tensor):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def clamp_code(clamp_mode) -> int:
    """Reference: `raise "Need to choose clamp mode"` is a TypeError in py3
    (generators/volumetric_rendering.py:41-46)."""
    if clamp_mode not in CLAMP_MODES:
        raise TypeError("Need to choose clamp mode")
    return CLAMP_MODES[clamp_mode]


def volume_to_channels_last(vol: torch.Tensor) -> torch.Tensor:
    """[B,C,D,H,W] -> [B,D,H,W,C] (cng_volume_to_channels_last).

    A volume the encoder already emits in ``torch.channels_last_3d`` memory format IS the kernel's layout: its
    storage is returned as a view and no kernel runs (the encoder emitting the volume in the
    gather kernel's layout)."""
    if (isinstance(vol, torch.Tensor) and vol.is_cuda and vol.dim() == 5 and vol.dtype in (torch.float32, torch.float16, torch.bfloat16)
            and not vol.is_contiguous() and vol.is_contiguous(memory_format=torch.channels_last_3d)):
        # (a 16-bit channels-last volume -- the encoder under autocast -- is widened in place of being re-laid: .float() keeps
        # the memory format, so this is one pass and still no layout kernel)
        return vol.float().permute(0, 2, 3, 4, 1)
    if isinstance(vol, torch.Tensor) and vol.is_cuda and vol.dtype == torch.float16 and vol.dim() == 5:
        vol = vol.contiguous()
        B, C, D, H, W = vol.shape
        out = torch.empty((B, D, H, W, C), dtype=torch.float32, device=vol.device)
        with torch.cuda.device(vol.device), _timed("cng_volume_to_channels_last"):
            _lib.call("cng_volume_f16_to_channels_last", _ptr(vol), _ptr(out), B, C, D, H, W, _stream(vol))
        _count()
        return out
    vol = _f32(vol, "volume")
    B, C, D, H, W = vol.shape
    out = torch.empty((B, D, H, W, C), dtype=torch.float32, device=vol.device)
    with torch.cuda.device(vol.device), _timed("cng_volume_to_channels_last"):
        _lib.call("cng_volume_to_channels_last", _ptr(vol), _ptr(out), B, C, D, H, W, _stream(vol))
    _count()
    return out


def raymarch_gather_coarse(vol_cl, cam2world, rays_d_cam, t_lin, u_jitter, img_w, img_h, want_points=False):
    """K1 coarse.  Returns feat[B,R,S,C], t[B,R,S], points[B,R,S,3] or None."""
    vol_cl = _f32(vol_cl, "vol_ndhwc")
    Bv, D, H, W, C = vol_cl.shape
    cam2world = _f32(cam2world, "cam2world")
    rays_d_cam = _f32(rays_d_cam, "rays_d_cam")
    t_lin = _f32(t_lin, "t_lin")
    S = t_lin.numel()
    R = img_w * img_h
    B = cam2world.shape[0]
    if cam2world.shape != (B, 4, 4) or Bv not in (1, B):
        raise ValueError(f"cam2world {tuple(cam2world.shape)} does not match volume batch {Bv}")
    stride = 0 if (Bv == 1 and B > 1) else C * D * H * W      # one object seen from B cameras shares its volume
    if u_jitter is not None:
        u_jitter = _f32(u_jitter, "u_jitter")
        if u_jitter.numel() != B * R * S:
            raise ValueError("u_jitter must hold B*R*S draws")
    dev = vol_cl.device
    feat = torch.empty((B, R, S, C), dtype=torch.float32, device=dev)
    t_out = torch.empty((B, R, S), dtype=torch.float32, device=dev)
    pts = torch.empty((B, R, S, 3), dtype=torch.float32, device=dev) if want_points else None
    with torch.cuda.device(dev), _timed("cng_raymarch_gather_coarse"):
        _lib.call("cng_raymarch_gather_coarse", _ptr(vol_cl), stride, B, C, D, H, W, _ptr(cam2world), _ptr(rays_d_cam),
                  _ptr(t_lin), _ptr(u_jitter), img_w, img_h, S, _ptr(feat), _ptr(t_out), _ptr(pts), _stream(vol_cl))
    _count()
    return feat, t_out, pts


def raymarch_gather_fine(vol_cl, cam2world, rays_d_cam, t_fine, img_w, img_h, want_points=False):
    """K1 fine.  t_fine [B,R,S].  Returns feat[B,R,S,C], points or None."""
    vol_cl = _f32(vol_cl, "vol_ndhwc")
    Bv, D, H, W, C = vol_cl.shape
    cam2world = _f32(cam2world, "cam2world")
    rays_d_cam = _f32(rays_d_cam, "rays_d_cam")
    t_fine = _f32(t_fine, "t_fine")
    R = img_w * img_h
    B = cam2world.shape[0]
    if Bv not in (1, B):
        raise ValueError(f"cam2world {tuple(cam2world.shape)} does not match volume batch {Bv}")
    stride = 0 if (Bv == 1 and B > 1) else C * D * H * W
    S = t_fine.numel() // (B * R)
    dev = vol_cl.device
    feat = torch.empty((B, R, S, C), dtype=torch.float32, device=dev)
    pts = torch.empty((B, R, S, 3), dtype=torch.float32, device=dev) if want_points else None
    with torch.cuda.device(dev), _timed("cng_raymarch_gather_fine"):
        _lib.call("cng_raymarch_gather_fine", _ptr(vol_cl), stride, B, C, D, H, W, _ptr(cam2world), _ptr(rays_d_cam),
                  _ptr(t_fine), img_w, img_h, S, _ptr(feat), _ptr(pts), _stream(vol_cl))
    _count()
    return feat, pts


_DUMMY_VOL = {}


def _dummy_volume(dev) -> torch.Tensor:
    """A 1x1x1x32 volume for K1's points-only mode (the position-input SIREN has no feature volume; the kernel never reads it)."""
    key = (dev.type, dev.index)
    v = _DUMMY_VOL.get(key)
    if v is None:
        v = _DUMMY_VOL[key] = torch.zeros((1, 1, 1, 1, 32), dtype=torch.float32, device=dev)
    return v


def raymarch_points_coarse(cam2world, rays_d_cam, t_lin, u_jitter, img_w, img_h):
    """K1 coarse in points-only mode (a1-a3 without a4).  Returns t[B,R,S], points[B,R,S,3]."""
    cam2world, rays_d_cam, t_lin = _f32(cam2world, "cam2world"), _f32(rays_d_cam, "rays_d_cam"), _f32(t_lin, "t_lin")
    B, S, R = cam2world.shape[0], t_lin.numel(), img_w * img_h
    dev = cam2world.device
    if u_jitter is not None:
        u_jitter = _f32(u_jitter, "u_jitter")
    vol = _dummy_volume(dev)
    t_out = torch.empty((B, R, S), dtype=torch.float32, device=dev)
    pts = torch.empty((B, R, S, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _timed("cng_raymarch_gather_coarse"):
        _lib.call("cng_raymarch_gather_coarse", _ptr(vol), 0, B, 32, 1, 1, 1, _ptr(cam2world), _ptr(rays_d_cam), _ptr(t_lin), _ptr(u_jitter),
                  img_w, img_h, S, None, _ptr(t_out), _ptr(pts), _stream(cam2world))
    _count()
    return t_out, pts


def raymarch_points_fine(cam2world, rays_d_cam, t_fine, img_w, img_h):
    """K1 fine in points-only mode (a10).  t_fine [B,R,S] -> points[B,R,S,3]."""
    cam2world, rays_d_cam, t_fine = _f32(cam2world, "cam2world"), _f32(rays_d_cam, "rays_d_cam"), _f32(t_fine, "t_fine")
    B, R = cam2world.shape[0], img_w * img_h
    S = t_fine.numel() // (B * R)
    dev = cam2world.device
    vol = _dummy_volume(dev)
    pts = torch.empty((B, R, S, 3), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev), _timed("cng_raymarch_gather_fine"):
        _lib.call("cng_raymarch_gather_fine", _ptr(vol), 0, B, 32, 1, 1, 1, _ptr(cam2world), _ptr(rays_d_cam), _ptr(t_fine), img_w, img_h, S,
                  None, _ptr(pts), _stream(cam2world))
    _count()
    return pts


def gather_points(vol_cl, points, want_index=False):
    """Trilinear lookup at caller-supplied world points [B,N,3] -> feat[B,N,C] (+ corner index)."""
    vol_cl = _f32(vol_cl, "vol_ndhwc")
    B, D, H, W, C = vol_cl.shape
    points = _f32(points, "points")
    if points.dim() != 3 or points.shape[0] != B or points.shape[2] != 3:
        raise ValueError(f"points must be [B={B}, N, 3], got {tuple(points.shape)}")
    N = points.shape[1]
    dev = vol_cl.device
    feat = torch.empty((B, N, C), dtype=torch.float32, device=dev)
    idx = torch.empty((B, N, 3), dtype=torch.int32, device=dev) if want_index else None
    with torch.cuda.device(dev), _timed("cng_gather_points"):
        _lib.call("cng_gather_points", _ptr(vol_cl), B, C, D, H, W, _ptr(points), N, _ptr(feat), _ptr(idx), _stream(vol_cl))
    _count()
    return (feat, idx) if want_index else feat


def film_parameters(global_feature, map_w, map_b) -> Tuple[torch.Tensor, torch.Tensor]:
    """a5: freq = 15 * linear(global)[:half] + 30, phase = linear(global)[half:]; each [B, n_out/2].  No autograd."""
    g, w, b = _f32(global_feature, "global_feature"), _f32(map_w, "map_w"), _f32(map_b, "map_b")
    B, z_dim = g.shape
    n_out = w.shape[0]
    if w.shape[1] != z_dim:
        raise ValueError(f"global feature has {z_dim} channels, the mapping network expects {w.shape[1]}")
    freq = torch.empty((B, n_out // 2), dtype=torch.float32, device=g.device)
    phase = torch.empty((B, n_out // 2), dtype=torch.float32, device=g.device)
    with torch.cuda.device(g.device), _timed("cng_film_parameters"):
        _lib.call("cng_film_parameters", _ptr(g), _ptr(w), _ptr(b), B, z_dim, n_out, _ptr(freq), _ptr(phase), _stream(g))
    _count()
    return freq, phase


_RES_SCRATCH = {}


def _res_scratch(dev) -> torch.Tensor:
    """Device buffer for the kept activations of residual blocks (cng_film_siren_res_scratch_bytes), one per device AND
    stream: launches on the same stream are ordered, launches on different streams must not share it."""
    with torch.cuda.device(dev):
        stream_id = torch.cuda.current_stream().cuda_stream
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(), stream_id)
    buf = _RES_SCRATCH.get(key)
    if buf is None:
        with torch.cuda.device(dev):
            n = int(_lib.load().cng_film_siren_res_scratch_bytes())
        buf = _RES_SCRATCH[key] = torch.empty((max(n, 16),), dtype=torch.uint8, device=dev)
    return buf


def film_siren_fwd(feat, layer_w: Sequence[torch.Tensor], layer_b: Sequence[torch.Tensor], freq, phase, final_w,
                   final_b, sigmoid_rgb: bool, precision: str = "bf16", res_save_mask: int = 0, res_add_mask: int = 0) -> torch.Tensor:
    """K2.  feat [B,N,C], freq/phase [B,L*HID] -> rgb_sigma [B,N,4].  ``res_*_mask``: residual blocks, see
    cng_film_siren_fwd_res in include/cng_b200.h."""
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    feat = _f32(feat, "feat")
    B, N, C = feat.shape
    L = len(layer_w)
    ws = [_f32(w, f"layer_w[{i}]") for i, w in enumerate(layer_w)]
    bs = [_f32(b, f"layer_b[{i}]") for i, b in enumerate(layer_b)]
    HID = ws[0].shape[0]
    freq, phase = _f32(freq, "freq"), _f32(phase, "phase")
    if freq.shape != (B, L * HID) or phase.shape != (B, L * HID):
        raise ValueError(f"freq/phase must be [B={B}, L*HID={L * HID}], got {tuple(freq.shape)} / {tuple(phase.shape)}")
    final_w, final_b = _f32(final_w, "final_w"), _f32(final_b, "final_b")
    dev = feat.device
    out = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    w_arr = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
    b_arr = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bs])
    code = PRECISIONS[precision]
    lib = _lib.load()
    ws_bytes = int(lib.cng_film_siren_workspace_bytes(B, C, HID, L, code))
    workspace = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev) if ws_bytes else None
    if res_save_mask or res_add_mask:
        scratch = _res_scratch(dev) if code != _lib.PREC_FP32 else None
        with torch.cuda.device(dev), _timed("cng_film_siren_fwd"):
            _lib.call("cng_film_siren_fwd_res", _ptr(feat), B, N, C, HID, L, w_arr, b_arr, _ptr(freq), _ptr(phase), _ptr(final_w),
                      _ptr(final_b), int(bool(sigmoid_rgb)), code, int(res_save_mask), int(res_add_mask), _ptr(workspace), ws_bytes,
                      _ptr(scratch), scratch.numel() if scratch is not None else 0, _ptr(out), _stream(feat))
        _count(1 if code == _lib.PREC_FP32 else 2)
        return out
    with torch.cuda.device(dev), _timed("cng_film_siren_fwd"):
        _lib.call("cng_film_siren_fwd", _ptr(feat), B, N, C, HID, L, w_arr, b_arr, _ptr(freq), _ptr(phase), _ptr(final_w),
                  _ptr(final_b), int(bool(sigmoid_rgb)), code, _ptr(workspace), ws_bytes, _ptr(out), _stream(feat))
    _count(1 if code == _lib.PREC_FP32 else 2)
    return out


def film_siren_fwd_gather(vol_cl, points, layer_w, layer_b, freq, phase, final_w, final_b, sigmoid_rgb: bool, precision: str = "bf16") -> torch.Tensor:
    """K2 with the trilinear lookup fused into its prologue (cng_film_siren_fwd_gather): vol_cl [B or 1, D,H,W,32], points [B,N,3]
    -> rgb_sigma [B,N,4]; the gathered features never go to HBM."""
    if precision not in ("bf16", "fp16"):
        raise ValueError("film_siren_fwd_gather: precision must be 'bf16' or 'fp16'")
    vol_cl, points = _f32(vol_cl, "vol_ndhwc"), _f32(points, "points")
    Bv, D, H, W, C = vol_cl.shape
    B, N = points.shape[0], points.shape[1]
    stride = 0 if (Bv == 1 and B > 1) else C * D * H * W
    L = len(layer_w)
    ws = [_f32(w, f"layer_w[{i}]") for i, w in enumerate(layer_w)]
    bs = [_f32(b, f"layer_b[{i}]") for i, b in enumerate(layer_b)]
    HID = ws[0].shape[0]
    freq, phase, final_w, final_b = _f32(freq, "freq"), _f32(phase, "phase"), _f32(final_w, "final_w"), _f32(final_b, "final_b")
    dev = vol_cl.device
    out = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    w_arr = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
    b_arr = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bs])
    code = PRECISIONS[precision]
    ws_bytes = int(_lib.load().cng_film_siren_workspace_bytes(B, C, HID, L, code))
    workspace = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev), _timed("cng_film_siren_fwd_gather"):
        _lib.call("cng_film_siren_fwd_gather", _ptr(vol_cl), stride, D, H, W, _ptr(points), B, N, C, HID, L, w_arr, b_arr, _ptr(freq), _ptr(phase),
                  _ptr(final_w), _ptr(final_b), int(bool(sigmoid_rgb)), code, _ptr(workspace), ws_bytes, _ptr(out), _stream(vol_cl))
    _count(2)
    return out


def composite_fwd(rgb_sigma, t, noise, noise_std: float, clamp_mode, white_back=False, last_back=False,
                  want_weights=True) -> Tuple[torch.Tensor, torch.Tensor, Optional[torch.Tensor]]:
    """K3.  rgb_sigma [..., S, 4], t [..., S(,1)] -> rgb [..., 3], dist [...], weights [..., S]."""
    code = clamp_code(clamp_mode)
    rgb_sigma = _f32(rgb_sigma, "rgb_sigma")
    S = rgb_sigma.shape[-2]
    lead = rgb_sigma.shape[:-2]
    n_rays = rgb_sigma.numel() // (4 * S)
    t = _f32(t, "t").reshape(n_rays, S)
    if noise is not None and noise_std != 0:
        noise = _f32(noise, "noise").reshape(n_rays, S)
    else:
        noise = None
    dev = rgb_sigma.device
    rgb = torch.empty((*lead, 3), dtype=torch.float32, device=dev)
    dist = torch.empty(tuple(lead), dtype=torch.float32, device=dev)
    weights = torch.empty((*lead, S), dtype=torch.float32, device=dev) if want_weights else None
    with torch.cuda.device(dev), _timed("cng_composite_fwd"):
        _lib.call("cng_composite_fwd", _ptr(rgb_sigma), _ptr(t), _ptr(noise), n_rays, S, float(noise_std), code,
                  int(bool(white_back)), int(bool(last_back)), _ptr(rgb), _ptr(dist), _ptr(weights), _stream(rgb_sigma))
    _count()
    return rgb, dist, weights


def sample_pdf(bins, weights, u, eps: float = 1e-5, want_inds=False):
    """K4.  bins [n,M+1], weights [n,M], u [n,K] -> samples [n,K] (+ int64 searchsorted indices)."""
    bins, weights, u = _f32(bins, "bins"), _f32(weights, "weights"), _f32(u, "u")
    n, M = weights.shape
    if bins.shape != (n, M + 1):
        raise ValueError(f"bins must be [n, M+1] = {(n, M + 1)}, got {tuple(bins.shape)}")
    K = u.shape[1]
    dev = bins.device
    samples = torch.empty((n, K), dtype=torch.float32, device=dev)
    inds = torch.empty((n, K), dtype=torch.int64, device=dev) if want_inds else None
    with torch.cuda.device(dev), _timed("cng_sample_pdf"):
        _lib.call("cng_sample_pdf", _ptr(bins), _ptr(weights), _ptr(u), n, M, K, float(eps), _ptr(samples), _ptr(inds), _stream(bins))
    _count()
    return (samples, inds) if want_inds else samples


def resample_from_coarse(t_coarse, weights, u, want_inds=False):
    """The sample_pdf call site of the generator fused: t_coarse, weights, u all [n,S] -> t_fine [n,S]."""
    t_coarse, weights, u = _f32(t_coarse, "t_coarse"), _f32(weights, "weights"), _f32(u, "u")
    S = t_coarse.shape[-1]
    n = t_coarse.numel() // S
    if weights.numel() != n * S or u.numel() != n * S:
        raise ValueError("t_coarse, weights and u must all hold n*S values")
    dev = t_coarse.device
    t_fine = torch.empty((n, S), dtype=torch.float32, device=dev)
    inds = torch.empty((n, S), dtype=torch.int64, device=dev) if want_inds else None
    with torch.cuda.device(dev), _timed("cng_resample_from_coarse"):
        _lib.call("cng_resample_from_coarse", _ptr(t_coarse), _ptr(weights), _ptr(u), n, S, _ptr(t_fine), _ptr(inds), _stream(t_coarse))
    _count()
    return (t_fine, inds) if want_inds else t_fine


def merge_composite(rgb_sigma_fine, rgb_sigma_coarse, t_fine, t_coarse, noise, rays_d_cam, B, img_h, img_w, noise_std,
                    clamp_mode, white_back=False, last_back=False, taps=False):
    """K3 final.  Returns pixels [B,3,H,W], depth [B,H,W] (+ dict of taps)."""
    code = clamp_code(clamp_mode)
    rgb_sigma_coarse = _f32(rgb_sigma_coarse, "rgb_sigma_coarse")
    t_coarse = _f32(t_coarse, "t_coarse")
    R = img_h * img_w
    S = rgb_sigma_coarse.numel() // (B * R * 4)
    two = rgb_sigma_fine is not None
    if two:
        rgb_sigma_fine, t_fine = _f32(rgb_sigma_fine, "rgb_sigma_fine"), _f32(t_fine, "t_fine")
    n = 2 * S if two else S
    noise = _f32(noise, "noise") if (noise is not None and noise_std != 0) else None
    rays_d_cam = _f32(rays_d_cam, "rays_d_cam")
    dev = rgb_sigma_coarse.device
    pixels = torch.empty((B, 3, img_h, img_w), dtype=torch.float32, device=dev)
    depth = torch.empty((B, img_h, img_w), dtype=torch.float32, device=dev)
    rgb = dist = order = None
    if taps:
        rgb = torch.empty((B, R, 3), dtype=torch.float32, device=dev)
        dist = torch.empty((B, R), dtype=torch.float32, device=dev)
        order = torch.empty((B, R, n), dtype=torch.int32, device=dev)
    with torch.cuda.device(dev), _timed("cng_merge_composite"):
        _lib.call("cng_merge_composite", _ptr(rgb_sigma_fine) if two else None, _ptr(rgb_sigma_coarse),
                  _ptr(t_fine) if two else None, _ptr(t_coarse), _ptr(noise), _ptr(rays_d_cam), B, R, S, float(noise_std),
                  code, int(bool(white_back)), int(bool(last_back)), _ptr(pixels), _ptr(depth), _ptr(rgb), _ptr(dist),
                  _ptr(order), _stream(rgb_sigma_coarse))
    _count()
    if taps:
        return pixels, depth, {"rgb": rgb, "dist": dist, "order": order}
    return pixels, depth


# --------------------------------------------------------------------------------------------------
# backward entry points
# --------------------------------------------------------------------------------------------------
def merge_composite_bwd(rgb_sigma_fine, rgb_sigma_coarse, t_fine, t_coarse, noise, rays_d_cam, d_pixels, d_depth, B, img_h,
                        img_w, noise_std, clamp_mode, white_back=False, last_back=False):
    """Backward of merge_composite.  Returns (d_rgb_sigma_fine or None, d_rgb_sigma_coarse), each [B*R, S, 4]."""
    code = clamp_code(clamp_mode)
    rgb_sigma_coarse, t_coarse = _f32(rgb_sigma_coarse, "rgb_sigma_coarse"), _f32(t_coarse, "t_coarse")
    R = img_h * img_w
    S = rgb_sigma_coarse.numel() // (B * R * 4)
    two = rgb_sigma_fine is not None
    if two:
        rgb_sigma_fine, t_fine = _f32(rgb_sigma_fine, "rgb_sigma_fine"), _f32(t_fine, "t_fine")
    noise = _f32(noise, "noise") if (noise is not None and noise_std != 0) else None
    rays_d_cam = _f32(rays_d_cam, "rays_d_cam")
    d_pixels = _f32(d_pixels, "d_pixels") if d_pixels is not None else None
    d_depth = _f32(d_depth, "d_depth") if d_depth is not None else None
    dev = rgb_sigma_coarse.device
    d_coarse = torch.empty((B * R, S, 4), dtype=torch.float32, device=dev)
    d_fine = torch.empty((B * R, S, 4), dtype=torch.float32, device=dev) if two else None
    with torch.cuda.device(dev), _timed("cng_merge_composite_bwd"):
        _lib.call("cng_merge_composite_bwd", _ptr(rgb_sigma_fine) if two else None, _ptr(rgb_sigma_coarse),
                  _ptr(t_fine) if two else None, _ptr(t_coarse), _ptr(noise), _ptr(rays_d_cam), _ptr(d_pixels), _ptr(d_depth),
                  B, R, S, float(noise_std), code, int(bool(white_back)), int(bool(last_back)), _ptr(d_fine), _ptr(d_coarse),
                  _stream(rgb_sigma_coarse))
    _count()
    return d_fine, d_coarse


def scatter_points(dvol_cl, points, dfeat) -> None:
    """dvol_cl [B,D,H,W,C] += trilinear-weighted dfeat [B,N,C] at points [B,N,3] (in place)."""
    if not (dvol_cl.is_cuda and dvol_cl.dtype == torch.float32 and dvol_cl.is_contiguous()):
        raise RuntimeError("scatter_points: dvol_cl must be a contiguous fp32 CUDA tensor (the rendering path has no CPU implementation)")
    B, D, H, W, C = dvol_cl.shape
    points, dfeat = _f32(points, "points"), _f32(dfeat, "dfeat")
    N = points.numel() // (3 * B)
    with torch.cuda.device(dvol_cl.device), _timed("cng_scatter_points"):
        _lib.call("cng_scatter_points", _ptr(dvol_cl), B, C, D, H, W, _ptr(points), N, _ptr(dfeat), _stream(dvol_cl))
    _count()


def volume_from_channels_last(vol_cl: torch.Tensor) -> torch.Tensor:
    """[B,D,H,W,C] -> [B,C,D,H,W]."""
    vol_cl = _f32(vol_cl, "vol_ndhwc")
    B, D, H, W, C = vol_cl.shape
    out = torch.empty((B, C, D, H, W), dtype=torch.float32, device=vol_cl.device)
    with torch.cuda.device(vol_cl.device), _timed("cng_volume_from_channels_last"):
        _lib.call("cng_volume_from_channels_last", _ptr(vol_cl), _ptr(out), B, C, D, H, W, _stream(vol_cl))
    _count()
    return out


TILE_POINTS = 128            # points per operand tile of the tcgen05 kernels
TILE_IMAGE_BYTES = 65536     # one tile-layer of x / dz (include/cng_b200.h, "Dump formats")


def g_image_bytes() -> int:
    """Bytes of one tile-layer of the cos(u) dump in the library's current format (cng_film_siren_g_dump_bits: fp16 -> 65536,
    8-bit codes -> 32768)."""
    return TILE_POINTS * 256 * int(_lib.load().cng_film_siren_g_dump_bits()) // 8

FEAT_IMAGE_BYTES = 16384


def film_siren_fwd_train(feat, layer_w, layer_b, freq, phase, final_w, final_b, sigmoid_rgb: bool, precision: str = "fp16",
                         res_save_mask: int = 0, res_add_mask: int = 0, dumps=None):
    """Training-mode K2 (the backward's recompute): rgb_sigma [B,N,4] plus the dumps in the formats of include/cng_b200.h --
    x [L,T,65536] uint8 (operand tile images), g = cos(u) [L,T,g_image_bytes()] uint8 (fp16 or 8-bit codes, epilogue order), feat [T,16384] uint8 -- with
    T = B * ceil(N / 128)."""
    if precision not in ("bf16", "fp16"):
        raise ValueError("film_siren_fwd_train: precision must be 'bf16' or 'fp16'")
    feat = _f32(feat, "feat")
    B, N, C = feat.shape
    L = len(layer_w)
    ws = [_f32(w, f"layer_w[{i}]") for i, w in enumerate(layer_w)]
    bs = [_f32(b, f"layer_b[{i}]") for i, b in enumerate(layer_b)]
    HID = ws[0].shape[0]
    freq, phase = _f32(freq, "freq"), _f32(phase, "phase")
    final_w, final_b = _f32(final_w, "final_w"), _f32(final_b, "final_b")
    dev = feat.device
    T = B * ((N + TILE_POINTS - 1) // TILE_POINTS)
    out = torch.empty((B, N, 4), dtype=torch.float32, device=dev)
    if dumps is not None:                      # caller-owned (pooled) dump buffers of exactly this shape
        xs, gs, fd = dumps
        if xs.shape != (L, T, TILE_IMAGE_BYTES) or gs.shape != (L, T, g_image_bytes()) or fd.shape != (T, FEAT_IMAGE_BYTES):
            raise ValueError("film_siren_fwd_train: dump buffers of the wrong shape")
    else:
        xs = torch.empty((L, T, TILE_IMAGE_BYTES), dtype=torch.uint8, device=dev)
        gs = torch.empty((L, T, g_image_bytes()), dtype=torch.uint8, device=dev)
        fd = torch.empty((T, FEAT_IMAGE_BYTES), dtype=torch.uint8, device=dev)
    w_arr = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
    b_arr = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bs])
    lib = _lib.load()
    code = PRECISIONS[precision]
    ws_bytes = int(lib.cng_film_siren_workspace_bytes(B, C, HID, L, code))
    workspace = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=dev)
    scratch = _res_scratch(dev) if (res_save_mask or res_add_mask) else None
    with torch.cuda.device(dev), _timed("cng_film_siren_fwd_train"):
        _lib.call("cng_film_siren_fwd_train", _ptr(feat), B, N, C, HID, L, w_arr, b_arr, _ptr(freq), _ptr(phase), _ptr(final_w),
                  _ptr(final_b), int(bool(sigmoid_rgb)), code, int(res_save_mask), int(res_add_mask), _ptr(workspace), ws_bytes,
                  _ptr(scratch), scratch.numel() if scratch is not None else 0, _ptr(out), _ptr(xs), _ptr(gs), _ptr(fd), _stream(feat))
    _count(2)
    return out, xs, gs, fd


def film_siren_wt_images(layer_w, final_w, freq=None) -> torch.Tensor:
    """Operand images of (diag(freq_l) W_l)^T (plain W_l^T without ``freq`` [L*HID]) and of the head for the dgrad chain."""
    L = len(layer_w)
    ws = [_f32(w, f"layer_w[{i}]") for i, w in enumerate(layer_w)]
    final_w = _f32(final_w, "final_w")
    dev = final_w.device
    lib = _lib.load()
    img = torch.empty((int(lib.cng_film_siren_wt_image_bytes(L)),), dtype=torch.uint8, device=dev)
    w_arr = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
    with torch.cuda.device(dev), _timed("cng_film_siren_wt_images"):
        _lib.call("cng_film_siren_wt_images", w_arr, _ptr(_f32(freq, "freq")) if freq is not None else None, _ptr(final_w), ws[0].shape[1],
                  ws[0].shape[0], L, _ptr(img), _stream(final_w))
    _count()
    return img


def _tile_ptr(dump: torch.Tensor, layer: int, tile: int, stride_tiles: int):
    """Address of tile ``tile`` of layer ``layer`` inside a [L, stride_tiles, bytes per tile] uint8 dump."""
    return ctypes.c_void_p(dump.data_ptr() + (layer * stride_tiles + tile) * dump.shape[2])


def film_siren_dgrad(d_out, out, sigmoid_rgb: bool, L: int, wt_images, g_dump, d_final_b_acc, res_save_mask: int = 0, res_add_mask: int = 0,
                     tile_offset: int = 0, d_feat=None):
    """The fused dgrad chain (cng_film_siren_dgrad): d_out / out [P,4] -> (d_feat [P,32], dz tile images [L,T,65536] uint8);
    accumulates the head bias gradient into d_final_b_acc [4].  ``g_dump`` [L, T_total, g_image_bytes()]: this call reads tiles
    [tile_offset, tile_offset + T) of every layer (a batch's dumps consumed item by item)."""
    if g_dump.dim() != 3 or g_dump.shape[2] != g_image_bytes():
        raise ValueError(f"film_siren_dgrad: g_dump must be [L, T, {g_image_bytes()}] uint8 (cng_film_siren_g_dump_bits)")
    d_out = _f32(d_out, "d_out")
    P = d_out.shape[0]
    out = _f32(out, "out") if out is not None else None
    dev = d_out.device
    T = (P + TILE_POINTS - 1) // TILE_POINTS
    dz = torch.empty((L, T, TILE_IMAGE_BYTES), dtype=torch.uint8, device=dev)
    if d_feat is None:
        d_feat = torch.empty((P, 32), dtype=torch.float32, device=dev)
    elif not (d_feat.is_cuda and d_feat.dtype == torch.float32 and d_feat.is_contiguous() and d_feat.shape == (P, 32)):
        raise RuntimeError("film_siren_dgrad: d_feat must be a contiguous float32 CUDA tensor [P, 32]")
    scratch = _res_scratch(dev) if (res_save_mask or res_add_mask) else None
    with torch.cuda.device(dev), _timed("cng_film_siren_dgrad"):
        _lib.call("cng_film_siren_dgrad", _ptr(d_out), _ptr(out), int(bool(sigmoid_rgb)), P, L, _ptr(wt_images),
                  _tile_ptr(g_dump, 0, tile_offset, g_dump.shape[1]), g_dump.shape[1], _ptr(dz),
                  _ptr(d_feat), _ptr(d_final_b_acc), int(res_save_mask), int(res_add_mask), _ptr(scratch),
                  scratch.numel() if scratch is not None else 0, _stream(d_out))
    _count()
    return d_feat, dz


def film_siren_wgrad(dz_dump, x_dump, feat_dump, P: int, L: int, x_is_fp16: bool, d_w_acc, colsum_acc, tile_offset: int = 0) -> None:
    """Weight gradients by split-K over the points (cng_film_siren_wgrad): d_w_acc[l] [256,K_l] += dz_l^T x_l, colsum_acc [L,256] +=
    column sums of dz_l.  ``x_dump`` [>= L-1, T_total, 65536] / ``feat_dump`` [T_total, 16384]: tiles from ``tile_offset`` on."""
    arr = (ctypes.c_void_p * L)(*[t.data_ptr() for t in d_w_acc])
    stride = x_dump.shape[1] if x_dump is not None else 0
    xp = _tile_ptr(x_dump, 0, tile_offset, stride) if x_dump is not None else None
    fp = ctypes.c_void_p(feat_dump.data_ptr() + tile_offset * FEAT_IMAGE_BYTES)
    with torch.cuda.device(dz_dump.device), _timed("cng_film_siren_wgrad"):
        _lib.call("cng_film_siren_wgrad", _ptr(dz_dump), xp, stride, fp, P, L, int(bool(x_is_fp16)), arr, _ptr(colsum_acc), _stream(dz_dump))
    _count()


def film_siren_head_wgrad(d_out, out, sigmoid_rgb: bool, x_dump, L: int, P: int, x_is_fp16: bool, d_final_w_acc, tile_offset: int = 0) -> None:
    """d_final_w_acc [4,256] += d_o^T x_L from the last layer's tile images of ``x_dump`` (cng_film_siren_head_wgrad)."""
    d_out = _f32(d_out, "d_out")
    out = _f32(out, "out") if out is not None else None
    with torch.cuda.device(d_out.device), _timed("cng_film_siren_head_wgrad"):
        _lib.call("cng_film_siren_head_wgrad", _ptr(d_out), _ptr(out), int(bool(sigmoid_rgb)), _tile_ptr(x_dump, L - 1, tile_offset, x_dump.shape[1]),
                  P, int(bool(x_is_fp16)), _ptr(d_final_w_acc), _stream(d_out))
    _count()


_BWD_WS = {}


def film_siren_bwd(feat, d_out, layer_w, layer_b, freq, phase, final_w, final_b, sigmoid_rgb: bool,
                   d_feat, d_w_acc, colsum_acc, d_final_w_acc, d_final_b_acc, res_save_mask: int = 0, res_add_mask: int = 0) -> None:
    """The MLP backward of one chunk of points of one item as ONE library call (cng_film_siren_bwd): feat [P,C], d_out [P,4];
    writes d_feat [P,C], accumulates into d_w_acc[l] (= dz'^T x), colsum_acc [L,HID] (= colsum dz'), both without the FiLM
    frequency (include/cng_b200.h), d_final_w_acc, d_final_b_acc (all fp32)."""
    P, C = feat.shape
    L = len(layer_w)
    HID = layer_w[0].shape[0]
    dev = feat.device
    for name, t in (("feat", feat), ("d_out", d_out), ("d_feat", d_feat), ("colsum_acc", colsum_acc), ("d_final_w_acc", d_final_w_acc),
                    ("d_final_b_acc", d_final_b_acc), ("freq", freq), ("phase", phase)):
        if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()):
            raise RuntimeError(f"film_siren_bwd: {name} must be a contiguous float32 CUDA tensor")
    lib = _lib.load()
    with torch.cuda.device(dev):
        need = int(lib.cng_film_siren_bwd_workspace_bytes(P, C, HID, L))
        if need == 0:
            raise _lib.CngError("cng_film_siren_bwd", -2, f"the MLP backward is built for C=32, HID=256, L<=16 (got C={C}, HID={HID}, L={L})")
        key = (dev.index, torch.cuda.current_stream().cuda_stream)
        ws = _BWD_WS.get(key)
        if ws is None or ws.numel() < need + 256:
            _BWD_WS.pop(key, None)
            ws = _BWD_WS[key] = torch.empty((need + 256,), dtype=torch.uint8, device=dev)
        base = (-ws.data_ptr()) % 256
        scratch = _res_scratch(dev) if (res_save_mask or res_add_mask) else None
        arr = lambda ts: (ctypes.c_void_p * len(ts))(*[t.data_ptr() for t in ts])
        with _timed("cng_film_siren_bwd"):
            _lib.call("cng_film_siren_bwd", _ptr(feat), _ptr(d_out), P, C, HID, L, arr(layer_w), arr(layer_b),
                      _ptr(freq), _ptr(phase), _ptr(final_w), _ptr(final_b), int(bool(sigmoid_rgb)), int(res_save_mask),
                      int(res_add_mask), ctypes.c_void_p(ws.data_ptr() + base), need, _ptr(scratch), scratch.numel() if scratch is not None else 0,
                      _ptr(d_feat), arr(d_w_acc), _ptr(colsum_acc), _ptr(d_final_w_acc), _ptr(d_final_b_acc), _stream(feat))
    _count(6)


_GN_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def group_norm_channels_last(x, num_groups: int, weight, bias, eps: float):
    """GroupNorm forward on a channels-last tensor (cng_group_norm_fwd).  x: [N, C, *spatial] whose memory is [N, *spatial, C]
    (channels_last / channels_last_3d), fp32 / fp16 / bf16.  Returns (y in x's dtype and format, mean [N,G], rstd [N,G])."""
    N, C = x.shape[0], x.shape[1]
    S = x.numel() // max(N * C, 1)
    dev = x.device
    y = torch.empty_like(x)               # preserves the channels-last strides
    mean = torch.empty((N, num_groups), dtype=torch.float32, device=dev)
    rstd = torch.empty((N, num_groups), dtype=torch.float32, device=dev)
    sums = torch.empty((N, num_groups, 2), dtype=torch.float64, device=dev)
    w = _f32(weight, "weight") if weight is not None else None
    b = _f32(bias, "bias") if bias is not None else None
    with torch.cuda.device(dev), _timed("cng_group_norm_fwd"):
        _lib.call("cng_group_norm_fwd", _ptr(x), _GN_DTYPES[x.dtype], N, S, C, num_groups, _ptr(w), _ptr(b), float(eps), _ptr(y), _ptr(mean),
                  _ptr(rstd), _ptr(sums), _stream(x))
    _count(2)
    return y, mean, rstd


def group_norm_channels_last_bwd(dy, x, num_groups: int, weight, mean, rstd):
    """GroupNorm backward (cng_group_norm_bwd): returns (dx in x's dtype / format, ds [N,C], db [N,C]) with ds = sum dy * xhat,
    db = sum dy per item and channel (d_weight = ds.sum(0), d_bias = db.sum(0))."""
    N, C = x.shape[0], x.shape[1]
    S = x.numel() // max(N * C, 1)
    dev = x.device
    dx = torch.empty_like(x)
    ds = torch.empty((N, C), dtype=torch.float32, device=dev)
    db = torch.empty((N, C), dtype=torch.float32, device=dev)
    w = _f32(weight, "weight") if weight is not None else None
    with torch.cuda.device(dev), _timed("cng_group_norm_bwd"):
        _lib.call("cng_group_norm_bwd", _ptr(dy), _ptr(x), _GN_DTYPES[x.dtype], N, S, C, num_groups, _ptr(w), _ptr(mean), _ptr(rstd), _ptr(dx),
                  _ptr(ds), _ptr(db), _stream(x))
    _count(2)
    return dx, ds, db


def merge_sort(t_fine, t_coarse, want_sorted: bool = False):
    """a11 alone: order [n_rays, 2S] int32 of the stable fine-first sort of cat([t_fine, t_coarse]) (and the sorted distances)."""
    t_fine, t_coarse = _f32(t_fine, "t_fine"), _f32(t_coarse, "t_coarse")
    S = t_coarse.shape[-1] if t_coarse.shape[-1] != 1 else t_coarse.shape[-2]
    n_rays = t_coarse.numel() // S
    tf, tc = t_fine.reshape(n_rays, S), t_coarse.reshape(n_rays, S)
    order = torch.empty((n_rays, 2 * S), dtype=torch.int32, device=tc.device)
    t_sorted = torch.empty((n_rays, 2 * S), dtype=torch.float32, device=tc.device) if want_sorted else None
    with torch.cuda.device(tc.device), _timed("cng_merge_sort"):
        _lib.call("cng_merge_sort", _ptr(tf), _ptr(tc), n_rays, S, _ptr(order), _ptr(t_sorted), _stream(tc))
    _count()
    return (order, t_sorted) if want_sorted else order


def composite_bwd(rgb_sigma, t, noise, d_rgb, d_dist, noise_std: float, clamp_mode, white_back=False, last_back=False) -> torch.Tensor:
    """Backward of composite_fwd w.r.t. rgb_sigma: d_rgb [..., 3] and / or d_dist [...] (None allowed) -> [..., S, 4]."""
    code = clamp_code(clamp_mode)
    rgb_sigma = _f32(rgb_sigma, "rgb_sigma")
    S = rgb_sigma.shape[-2]
    n_rays = rgb_sigma.numel() // (4 * S)
    t = _f32(t, "t").reshape(n_rays, S)
    noise = _f32(noise, "noise").reshape(n_rays, S) if (noise is not None and noise_std != 0) else None
    d_rgb = _f32(d_rgb, "d_rgb").reshape(n_rays, 3) if d_rgb is not None else None
    d_dist = _f32(d_dist, "d_dist").reshape(n_rays) if d_dist is not None else None
    out = torch.empty_like(rgb_sigma)
    with torch.cuda.device(rgb_sigma.device), _timed("cng_composite_bwd"):
        _lib.call("cng_composite_bwd", _ptr(rgb_sigma), _ptr(t), _ptr(noise), _ptr(d_rgb), _ptr(d_dist), n_rays, S, float(noise_std), code,
                  int(bool(white_back)), int(bool(last_back)), _ptr(out), _stream(rgb_sigma))
    _count()
    return out


def render_fwd(vol_cl, cam2world, rays_d_cam, t_lin, layer_w, layer_b, freq, phase, final_w, final_b, sigmoid_rgb, precision,
               u_jitter, noise_coarse, u_resample, noise_final, img_w, img_h, hierarchical, noise_std, clamp_mode,
               white_back=False, last_back=False):
    """The whole forward in one C-ABI call (cng_render_fwd).  Returns pixels [B,3,H,W], depth [B,H,W]."""
    if precision not in PRECISIONS:
        raise ValueError(f"precision must be one of {sorted(PRECISIONS)}, got {precision!r}")
    code = clamp_code(clamp_mode)
    vol_cl = _f32(vol_cl, "vol_ndhwc")
    Bv, D, H, W, C = vol_cl.shape
    cam2world, rays_d_cam, t_lin = _f32(cam2world, "cam2world"), _f32(rays_d_cam, "rays_d_cam"), _f32(t_lin, "t_lin")
    B, S, R = cam2world.shape[0], t_lin.numel(), img_w * img_h
    if Bv not in (1, B):
        raise ValueError(f"cam2world {tuple(cam2world.shape)} does not match volume batch {Bv}")
    stride = 0 if (Bv == 1 and B > 1) else C * D * H * W
    L = len(layer_w)
    ws = [_f32(w, f"layer_w[{i}]") for i, w in enumerate(layer_w)]
    bs = [_f32(b, f"layer_b[{i}]") for i, b in enumerate(layer_b)]
    HID = ws[0].shape[0]
    freq, phase, final_w, final_b = _f32(freq, "freq"), _f32(phase, "phase"), _f32(final_w, "final_w"), _f32(final_b, "final_b")
    if freq.shape != (B, L * HID) or phase.shape != (B, L * HID):
        raise ValueError(f"freq/phase must be [B={B}, L*HID={L * HID}]")
    u_jitter = _f32(u_jitter, "u_jitter") if u_jitter is not None else None
    use_noise = noise_std != 0
    noise_coarse = _f32(noise_coarse, "noise_coarse") if (use_noise and noise_coarse is not None) else None
    noise_final = _f32(noise_final, "noise_final") if (use_noise and noise_final is not None) else None
    u_resample = _f32(u_resample, "u_resample") if hierarchical else None
    for name, t, n in (("u_jitter", u_jitter, B * R * S), ("noise_coarse", noise_coarse, B * R * S), ("u_resample", u_resample, B * R * S),
                       ("noise_final", noise_final, B * R * S * (2 if hierarchical else 1))):
        if t is not None and t.numel() != n:
            raise ValueError(f"{name} must hold {n} draws, got {t.numel()}")
    dev = vol_cl.device
    pcode = PRECISIONS[precision]
    lib = _lib.load()
    ws_bytes = int(lib.cng_render_workspace_bytes(B, img_w, img_h, S, C, HID, L, int(bool(hierarchical)), pcode))
    workspace = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    pixels = torch.empty((B, 3, img_h, img_w), dtype=torch.float32, device=dev)
    depth = torch.empty((B, img_h, img_w), dtype=torch.float32, device=dev)
    w_arr = (ctypes.c_void_p * L)(*[w.data_ptr() for w in ws])
    b_arr = (ctypes.c_void_p * L)(*[b.data_ptr() for b in bs])
    with torch.cuda.device(dev), _timed("cng_render_fwd"):
        _lib.call("cng_render_fwd", _ptr(vol_cl), stride, B, C, D, H, W, _ptr(cam2world), _ptr(rays_d_cam), _ptr(t_lin), img_w, img_h, S,
                  HID, L, w_arr, b_arr, _ptr(freq), _ptr(phase), _ptr(final_w), _ptr(final_b), int(bool(sigmoid_rgb)), pcode,
                  _ptr(u_jitter), _ptr(noise_coarse), _ptr(u_resample), _ptr(noise_final), int(bool(hierarchical)), float(noise_std), code,
                  int(bool(white_back)), int(bool(last_back)), _ptr(workspace), ws_bytes, _ptr(pixels), _ptr(depth), _stream(vol_cl))
    n_mlp = 1 if pcode == _lib.PREC_FP32 else 2
    _count((2 + 2 * n_mlp + 3) if hierarchical else (2 + n_mlp))
    return pixels, depth
