"""Autograd bridge of the rendering path.

The forward kernels stay the no-grad ones (K1, K2, K3', nothing saved per layer); each stage is a
``torch.autograd.Function`` whose backward launches the kernels of csrc/backward.cu:

    _ToChannelsLast   NCDHW <-> NDHWC                               cng_volume_{to,from}_channels_last
    _Gather*          d feat -> d volume (scatter-add)               cng_scatter_points
    _FilmSiren        d rgb_sigma -> d feat, d W/b, d freq/phase     cng_film_siren_bwd per chunk: activations
                      recomputed by the training-mode K2 (the fused tcgen05 forward that also writes
                      x_{l+1} tile images and g_l = cos(u_l)), then the tcgen05 dgrad chain
                      (dz = dy*g formed in its epilogue) and the tcgen05 split-K weight gradient --
                      no library GEMM (csrc/film_siren_bwd_tc.cu)
    _MergeComposite   d pixels, d depth -> d rgb_sigma (fine, coarse)  cng_merge_composite_bwd

Sample positions, distances and the coarse weights used for resampling carry no gradient, exactly
as in the reference (``torch.no_grad()`` blocks at generators/generators.py:57 and :111); the FiLM
mapping network (siren.py:550-553, a [B,256] x [256,4096] Linear) stays a differentiable torch op.
"""
from __future__ import annotations

from typing import List

import torch

from .. import ops
from .volumetric_rendering import camera_tables

import os

HAS_BACKWARD = True
CHUNK_ROWS = 1 << 20          # points per recompute chunk of the MLP backward (~13 GB of x / g / dz dumps at L = 8)
# "auto": when the x / g dumps of the WHOLE batch fit comfortably in free device memory (small per-GPU batches: the 8-GPU
# operating point), the forward of a training step runs the training-mode kernel once and keeps its dumps for the backward,
# which then skips the recompute; otherwise (and with "0") the forward keeps nothing and the backward recomputes per chunk.
KEEP_DUMPS = os.environ.get("CNG_KEEP_DUMPS", "auto")
KEEP_DUMPS_MEMORY_FRACTION = 0.35     # of the memory that is free (plus cached by the allocator) when the forward runs


class _DumpBuffers:
    """x / g / feature dump buffers of one kept forward, taken from a per-device pool and handed back after the backward (or when
    the graph is dropped without one).  The pool keeps them allocated across steps: tens of GB going through the caching
    allocator every step fragment it into cudaMalloc / cudaFree cycles, which synchronise the device (measured: the kept path
    was 19 % SLOWER than recomputing inside a full GAN step until the buffers became persistent)."""
    _pool = {}

    def __init__(self, L: int, tiles: int, dev):
        self.key = (dev.index if dev.index is not None else torch.cuda.current_device(), L, tiles, ops.g_image_bytes())
        free = self._pool.setdefault(self.key, [])
        if free:
            self.tensors = free.pop()
        else:
            self.tensors = (torch.empty((L, tiles, ops.TILE_IMAGE_BYTES), dtype=torch.uint8, device=dev),
                            torch.empty((L, tiles, ops.g_image_bytes()), dtype=torch.uint8, device=dev),
                            torch.empty((tiles, ops.FEAT_IMAGE_BYTES), dtype=torch.uint8, device=dev))

    def release(self) -> None:
        if self.tensors is not None:
            self._pool[self.key].append(self.tensors)
            self.tensors = None

    def __del__(self):
        self.release()

    @classmethod
    def pooled_bytes(cls, dev_index: int) -> int:
        return sum(sum(t.numel() for t in ts) for k, free in cls._pool.items() if k[0] == dev_index for ts in free)

    @classmethod
    def clear(cls) -> None:
        cls._pool.clear()


def _kept_items(B: int, N: int, L: int, dev) -> int:
    """How many of the B items run the training-mode kernel as their forward and keep its dumps for the backward (the first k; the
    others are recomputed chunk by chunk in the backward): as many as fit in KEEP_DUMPS_MEMORY_FRACTION of the free memory --
    all of them at the small per-GPU batches of multi-GPU training, a part of the batch at 16-32 images per GPU."""
    if KEEP_DUMPS == "0" or N > CHUNK_ROWS:
        return 0
    if KEEP_DUMPS == "1":
        return B
    if KEEP_DUMPS.startswith("first:"):                             # tests: exactly the first k items
        return min(B, int(KEEP_DUMPS[6:]))
    tpi = (N + ops.TILE_POINTS - 1) // ops.TILE_POINTS
    di = dev.index if dev.index is not None else torch.cuda.current_device()
    # a pooled buffer of a previous step is waiting: no new memory needed (and the choice is the same every step)
    pooled = [key[2] // tpi for key, free in _DumpBuffers._pool.items()
              if free and key[0] == di and key[1] == L and key[3] == ops.g_image_bytes() and key[2] % tpi == 0 and 0 < key[2] // tpi <= B]
    if pooled:
        return max(pooled)
    per_item = L * tpi * (ops.TILE_IMAGE_BYTES + ops.g_image_bytes()) + tpi * ops.FEAT_IMAGE_BYTES
    dz_item = L * tpi * ops.TILE_IMAGE_BYTES                        # one item's dz dump at backward time
    free, _ = torch.cuda.mem_get_info(dev)
    cached = torch.cuda.memory_reserved(dev) - torch.cuda.memory_allocated(dev)
    budget = KEEP_DUMPS_MEMORY_FRACTION * (free + cached) - dz_item
    return int(max(0, min(B, budget // per_item)))


class _ToChannelsLast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, volume):
        out = ops.volume_to_channels_last(volume)
        ctx.was_view = out.data_ptr() == volume.data_ptr()       # channels_last_3d input: no copy was made
        return out.contiguous() if ctx.was_view else out

    @staticmethod
    def backward(ctx, d_cl):
        d_cl = d_cl.contiguous()
        if ctx.was_view:                                          # gradient in the encoder's own (channels-last) format
            return d_cl.permute(0, 4, 1, 2, 3)
        return ops.volume_from_channels_last(d_cl)


def _scatter(vol_shape, points, d_feat):
    dvol = torch.zeros(vol_shape, dtype=torch.float32, device=d_feat.device)
    ops.scatter_points(dvol, points, d_feat.contiguous())
    return dvol


class _GatherCoarse(torch.autograd.Function):
    """K1 coarse; differentiable w.r.t. the (channels-last) volume only."""

    @staticmethod
    def forward(ctx, vol_cl, cam2world, rays_d_cam, t_lin, u_jitter, img_size):
        feat, t, pts = ops.raymarch_gather_coarse(vol_cl, cam2world, rays_d_cam, t_lin, u_jitter, img_size, img_size, want_points=True)
        ctx.save_for_backward(pts)
        ctx.vol_shape = vol_cl.shape
        ctx.mark_non_differentiable(t)
        return feat, t

    @staticmethod
    def backward(ctx, d_feat, _d_t):
        (pts,) = ctx.saved_tensors
        return _scatter(ctx.vol_shape, pts, d_feat), None, None, None, None, None


class _GatherFine(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol_cl, cam2world, rays_d_cam, t_fine, img_size):
        feat, pts = ops.raymarch_gather_fine(vol_cl, cam2world, rays_d_cam, t_fine, img_size, img_size, want_points=True)
        ctx.save_for_backward(pts)
        ctx.vol_shape = vol_cl.shape
        return feat

    @staticmethod
    def backward(ctx, d_feat):
        (pts,) = ctx.saved_tensors
        return _scatter(ctx.vol_shape, pts, d_feat), None, None, None, None


class _GatherPoints(torch.autograd.Function):
    @staticmethod
    def forward(ctx, vol_cl, points):
        ctx.save_for_backward(points)
        ctx.vol_shape = vol_cl.shape
        return ops.gather_points(vol_cl, points)

    @staticmethod
    def backward(ctx, d_feat):
        (points,) = ctx.saved_tensors
        return _scatter(ctx.vol_shape, points, d_feat), None


class _FilmSiren(torch.autograd.Function):
    """K2 forward (fused, nothing saved per layer); backward recomputes the activations chunk by chunk.

    The recompute always uses fp16 operands (11-bit significands), whatever ``precision`` the forward ran with (bf16 by
    default, or the exact fp32 kernel): the gradients are then those of a function within 4e-4 of the fp32 network -- closer
    to it than the bf16 forward itself (2.9e-3) -- and stay within the 2e-2 relative-L2 bar against fp32 autograd for every
    class (measured <= 0.9 %).  A bf16 recompute would be consistent with a bf16 forward bit for bit but doubles that error."""

    @staticmethod
    def forward(ctx, feat, freq, phase, final_w, final_b, sigmoid_rgb, precision, res_save, res_add, *wb):
        L = len(wb) // 2
        ws, bs = list(wb[:L]), list(wb[L:])
        if feat.shape[-1] != 32 or final_w.shape[1] != 256:
            # fail here, not deep inside backward(): the tcgen05 backward kernels are built for the shipped shape only
            raise NotImplementedError(f"training needs input_dim=32 and hidden_dim=256 (got {feat.shape[-1]}, {final_w.shape[1]}): "
                                      "the MLP backward (cng_film_siren_bwd) is built for that shape")
        B, N = feat.shape[0], feat.shape[1]
        ctx.dumps, ctx.kept = None, 0
        k = _kept_items(B, N, L, feat.device) if precision != "fp32" else 0
        if k > 0:
            # the training-mode kernel IS the forward of the first k items (fp16 operands): its dumps stay alive until backward,
            # no recompute there; the other items take the inference kernel
            ctx.kept = k
            ctx.dumps = _DumpBuffers(L, k * ((N + ops.TILE_POINTS - 1) // ops.TILE_POINTS), feat.device)
            out = ops.film_siren_fwd_train(feat[:k], ws, bs, freq[:k], phase[:k], final_w, final_b, sigmoid_rgb, "fp16", res_save, res_add,
                                           dumps=ctx.dumps.tensors)[0]
            if k < B:
                rest = ops.film_siren_fwd(feat[k:], ws, bs, freq[k:], phase[k:], final_w, final_b, sigmoid_rgb, precision, res_save, res_add)
                out = torch.cat([out, rest], dim=0)
        else:
            out = ops.film_siren_fwd(feat, ws, bs, freq, phase, final_w, final_b, sigmoid_rgb, precision, res_save, res_add)
        ctx.save_for_backward(feat, freq, phase, final_w, final_b, out, *wb)
        ctx.sigmoid_rgb, ctx.L, ctx.res_save, ctx.res_add = sigmoid_rgb, L, res_save, res_add
        return out

    @staticmethod
    def backward(ctx, d_out):
        """Chunked: per item and chunk of points one cng_film_siren_bwd call (recompute with dumps, dgrad chain, split-K
        weight gradient, head -- all tcgen05, csrc/film_siren_bwd_tc.cu) accumulates dW'_l = dz'_l^T x_l and the column sums
        of dz'_l = dy_l * cos(u_l) (the FiLM frequency is folded into the dgrad operand, never applied elementwise); from them,
        per item:  dW = freq * dW',  db = freq * colsum',  dphase = colsum',  dfreq = rowsum(W * dW') + b * colsum'.
        Residual blocks (``res_add`` / ``res_save`` masks, siren.py:218-230): g_l is taken at the pre-activation that includes
        the re-added block input, so the per-layer rule is unchanged; the kept activation (output of the last ``save`` layer
        before l) additionally receives dz_l of the adding layer (inside the dgrad kernel)."""
        feat, freq, phase, final_w, final_b, out, *wb = ctx.saved_tensors
        L, H = ctx.L, final_w.shape[1]
        B, N, C = feat.shape
        if C != 32 or H != 256:
            raise NotImplementedError(f"the MLP backward is built for input_dim=32, hidden_dim=256 (got {C}, {H})")
        ws, bs = [w.detach().float().contiguous() for w in wb[:L]], [b.detach().float().contiguous() for b in wb[L:]]
        fw, fb = final_w.detach().float().contiguous(), final_b.detach().float().contiguous()
        dev = feat.device
        d_out = d_out.contiguous().float()
        d_feat = torch.empty_like(feat)
        d_fw = torch.zeros_like(fw)
        d_fb = torch.zeros((4,), dtype=torch.float32, device=dev)
        # per item: dW' and the column sums of dz' (they are turned into dW / db / dphase / dfreq with the item's freq below,
        # batched over the items: a handful of launches per backward instead of a handful per item)
        dW_all = [torch.zeros((B,) + tuple(w.shape), dtype=torch.float32, device=dev) for w in ws]
        colsum_all = torch.zeros((B, L, H), dtype=torch.float32, device=dev)
        fr = freq.detach().float().contiguous()
        ph = phase.detach().float().contiguous()
        dumps, ctx.dumps = ctx.dumps, None
        tiles_per_item = (N + ops.TILE_POINTS - 1) // ops.TILE_POINTS
        for b in range(B):
            dW_item = [d[b] for d in dW_all]
            if dumps is not None and b < ctx.kept:
                # kept dumps (one training-mode forward for the first ctx.kept items): dgrad chain, weight gradient and head per item,
                # each reading its own tiles of every layer
                xs, gs, fd = dumps.tensors
                wt = ops.film_siren_wt_images(ws, fw, fr[b])
                _, dz = ops.film_siren_dgrad(d_out[b], out[b], ctx.sigmoid_rgb, L, wt, gs, d_fb, ctx.res_save, ctx.res_add,
                                             tile_offset=b * tiles_per_item, d_feat=d_feat[b])
                ops.film_siren_wgrad(dz, xs, fd, N, L, True, dW_item, colsum_all[b], tile_offset=b * tiles_per_item)
                ops.film_siren_head_wgrad(d_out[b], out[b], ctx.sigmoid_rgb, xs, L, N, True, d_fw, tile_offset=b * tiles_per_item)
                del dz
                continue
            for r0 in range(0, N, CHUNK_ROWS):
                r1 = min(N, r0 + CHUNK_ROWS)
                ops.film_siren_bwd(feat[b, r0:r1].detach().contiguous(), d_out[b, r0:r1].contiguous(), ws, bs, fr[b], ph[b], fw, fb,
                                   ctx.sigmoid_rgb, d_feat[b, r0:r1], dW_item, colsum_all[b], d_fw, d_fb, ctx.res_save, ctx.res_add)
        if dumps is not None:
            dumps.release()                               # stream-ordered: the next forward's writes queue behind this backward's reads
        frv = fr.view(B, L, H)
        d_phase = colsum_all.reshape(B, L * H)
        wdw = torch.stack([(ws[l].unsqueeze(0) * dW_all[l]).sum(2) for l in range(L)], dim=1)       # [B, L, H]
        d_freq = (wdw + torch.stack(bs).unsqueeze(0) * colsum_all).reshape(B, L * H)
        d_ws = [(dW_all[l] * frv[:, l, :, None]).sum(0) for l in range(L)]
        d_bs = [(colsum_all[:, l] * frv[:, l]).sum(0) for l in range(L)]
        return (d_feat, d_freq, d_phase, d_fw.to(final_w.dtype), d_fb.to(final_b.dtype), None, None, None, None,
                *[g.to(p.dtype) for g, p in zip(d_ws, wb[:L])], *[g.to(p.dtype) for g, p in zip(d_bs, wb[L:])])


class _MergeComposite(torch.autograd.Function):
    @staticmethod
    def forward(ctx, fine, coarse, t_fine, t_coarse, noise, rays_d_cam, B, img_size, noise_std, clamp_mode, white_back, last_back):
        pixels, depth = ops.merge_composite(fine, coarse, t_fine, t_coarse, noise, rays_d_cam, B, img_size, img_size, noise_std,
                                            clamp_mode, white_back, last_back)
        ctx.save_for_backward(fine, coarse, t_fine, t_coarse, noise, rays_d_cam)
        ctx.cfg = (B, img_size, noise_std, clamp_mode, white_back, last_back)
        return pixels, depth

    @staticmethod
    def backward(ctx, d_pixels, d_depth):
        fine, coarse, t_fine, t_coarse, noise, rays_d_cam = ctx.saved_tensors
        B, img_size, noise_std, clamp_mode, white_back, last_back = ctx.cfg
        d_fine, d_coarse = ops.merge_composite_bwd(fine, coarse, t_fine, t_coarse, noise, rays_d_cam,
                                                   d_pixels.contiguous() if d_pixels is not None else None,
                                                   d_depth.contiguous() if d_depth is not None else None,
                                                   B, img_size, img_size, noise_std, clamp_mode, white_back, last_back)
        if fine is not None:
            d_fine = d_fine.view_as(fine)
        return d_fine, d_coarse.view_as(coarse), None, None, None, None, None, None, None, None, None, None


def _mlp(net, feat, freq, phase):
    net.check_dropout()
    ws, bs = net.layer_parameters()
    return _FilmSiren.apply(feat, freq, phase, net.final_layer.weight, net.final_layer.bias, net.sigmoid_rgb, net.precision,
                            net.res_save_mask, net.res_add_mask, *ws, *bs)


def render_with_grad(gen, volume, global_feature, cam2worlds, img_size, fov, ray_start, ray_end, num_steps,
                     hierarchical_sample, kwargs):
    """generators/generators.py:33-187 with gradients to the SIREN parameters, the feature volume and the
    global feature.  Same launch sequence as ``ImplicitGenerator3d._render``."""
    clamp_mode, nerf_noise = kwargs["clamp_mode"], kwargs["nerf_noise"]
    white_back, last_back = kwargs.get("white_back", False), kwargs.get("last_back", False)
    ops.clamp_code(clamp_mode)
    draws = kwargs.get("draws") or {}
    net = gen.siren
    B, S, R = cam2worlds.shape[0], int(num_steps), int(img_size) ** 2
    dev = cam2worlds.device
    rays_d_cam, t_lin = camera_tables((img_size, img_size), S, fov, ray_start, ray_end, dev)

    def draw(name, fn, shape):
        t = draws.get(name)
        return fn(shape, device=dev) if t is None else t.to(dev)

    freq, phase = net.film_parameters(global_feature, B, dev)
    u_jitter = draw("u_jitter", torch.rand, (B, R, S, 1))
    if net.latent:
        # position-input SIREN (siren.py:1172-1224): no volume; the sample positions (no gradient, generators.py:57) are the operand rows
        vol_cl, C = None, 32
        with torch.no_grad():
            t_c, pts_c = ops.raymarch_points_coarse(cam2worlds, rays_d_cam, t_lin, u_jitter, img_size, img_size)
            feat_c = net.point_features(pts_c.view(B, R * S, 3))
    else:
        vol_cl = _ToChannelsLast.apply(volume.float())
        C = vol_cl.shape[-1]
        feat_c, t_c = _GatherCoarse.apply(vol_cl, cam2worlds, rays_d_cam, t_lin, u_jitter, img_size)
    coarse = _mlp(net, feat_c.view(B, R * S, C), freq, phase)
    if hierarchical_sample:
        with torch.no_grad():
            noise_c = draw("noise_coarse", torch.randn, (B, R, S, 1))
            _, _, w_c = ops.composite_fwd(coarse.detach().view(B, R, S, 4), t_c, noise_c, nerf_noise, clamp_mode)
            u_re = draw("u_resample", torch.rand, (B * R, S))
            t_f = ops.resample_from_coarse(t_c, w_c, u_re)
        if net.latent:
            with torch.no_grad():
                feat_f = net.point_features(ops.raymarch_points_fine(cam2worlds, rays_d_cam, t_f, img_size, img_size).view(B, R * S, 3))
        else:
            feat_f = _GatherFine.apply(vol_cl, cam2worlds, rays_d_cam, t_f, img_size)
        fine = _mlp(net, feat_f.view(B, R * S, C), freq, phase)
        noise_f = draw("noise_final", torch.randn, (B, R, 2 * S, 1))
        return _MergeComposite.apply(fine, coarse, t_f, t_c, noise_f, rays_d_cam, B, img_size, nerf_noise, clamp_mode,
                                     white_back, last_back)
    noise_f = draw("noise_final" if "noise_final" in draws else "noise_coarse", torch.randn, (B, R, S, 1))
    return _MergeComposite.apply(None, coarse, None, t_c, noise_f, rays_d_cam, B, img_size, nerf_noise, clamp_mode,
                                 white_back, last_back)


def render_library(gen, z, cam2worlds, img_size, fov, ray_start, ray_end, num_steps, hierarchical_sample, kwargs):
    """generators/generators.py:33-187 for the decoders whose MLP is a PyTorch op chain (generators/siren_library.py): K1 in
    points-only mode, ``gen.siren(points, z, ...)`` (library MLP around the library's trilinear kernels, differentiable when grad
    is enabled), K3 / K4 for the resampling (no gradient, generators.py:111), K3' (with its backward kernel when grad is enabled)."""
    clamp_mode, nerf_noise = kwargs["clamp_mode"], kwargs["nerf_noise"]
    white_back, last_back = kwargs.get("white_back", False), kwargs.get("last_back", False)
    ops.clamp_code(clamp_mode)
    draws = kwargs.get("draws") or {}
    net = gen.siren
    B, S, R = cam2worlds.shape[0], int(num_steps), int(img_size) ** 2
    dev = cam2worlds.device
    rays_d_cam, t_lin = camera_tables((img_size, img_size), S, fov, ray_start, ray_end, dev)
    grad = torch.is_grad_enabled()

    def draw(name, fn, shape):
        t = draws.get(name)
        return fn(shape, device=dev) if t is None else t.to(dev)

    def composite(fine, coarse, t_f, t_c, noise):
        if grad and (coarse.requires_grad or (fine is not None and fine.requires_grad)):
            return _MergeComposite.apply(fine, coarse, t_f, t_c, noise, rays_d_cam, B, img_size, nerf_noise, clamp_mode, white_back, last_back)
        return ops.merge_composite(fine, coarse, t_f, t_c, noise, rays_d_cam, B, img_size, img_size, nerf_noise, clamp_mode, white_back, last_back)

    u_jitter = draw("u_jitter", torch.rand, (B, R, S, 1))
    with torch.no_grad():
        t_c, pts_c = ops.raymarch_points_coarse(cam2worlds, rays_d_cam, t_lin, u_jitter, img_size, img_size)
    coarse = net(pts_c.view(B, R * S, 3), z, img_size, S).float().contiguous()
    if hierarchical_sample:
        with torch.no_grad():
            noise_c = draw("noise_coarse", torch.randn, (B, R, S, 1))
            _, _, w_c = ops.composite_fwd(coarse.detach().view(B, R, S, 4), t_c, noise_c, nerf_noise, clamp_mode)
            u_re = draw("u_resample", torch.rand, (B * R, S))
            t_f = ops.resample_from_coarse(t_c, w_c, u_re)
            pts_f = ops.raymarch_points_fine(cam2worlds, rays_d_cam, t_f, img_size, img_size)
        fine = net(pts_f.view(B, R * S, 3), z, img_size, S).float().contiguous()
        noise_f = draw("noise_final", torch.randn, (B, R, 2 * S, 1))
        return composite(fine, coarse, t_f, t_c, noise_f)
    noise_f = draw("noise_final" if "noise_final" in draws else "noise_coarse", torch.randn, (B, R, S, 1))
    return composite(None, coarse, None, t_c, noise_f)


def siren_forward_with_grad(net, points, volume, global_feature):
    """``siren(points, z, img_size, num_steps)`` with gradients (siren.py:540-580)."""
    vol_cl = _ToChannelsLast.apply(volume.float())
    freq, phase = net.film_parameters(global_feature, volume.shape[0], volume.device)
    feat = _GatherPoints.apply(vol_cl, points.detach())
    return _mlp(net, feat, freq, phase)
