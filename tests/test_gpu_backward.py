"""Parity of the backward kernels against torch autograd through the CPU oracle.
B200 only (-m gpu).  fp32 stages: <= 1e-4 relative; the MLP backward recomputes with bf16 GEMM operands
(like the forward tensor-core path), so its gradients are compared by relative L2 error and cosine."""
import numpy as np
import pytest
import torch

from conftest import fixture_inputs
from oracle import nerf_path as oracle

pytestmark = pytest.mark.gpu
FOV = 49.134342641202636


@pytest.fixture(scope="module")
def ops():
    from conditioned_nerf_gan_b200 import _lib, ops as _ops
    _lib.load()
    return _ops


def dev(t):
    return t.to("cuda")


def rel_l2(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def cosine(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a @ b) / (a.norm() * b.norm() + 1e-30))


@pytest.mark.parametrize("B,img,S,two", [(2, 6, 8, True), (1, 8, 24, True), (2, 4, 48, True), (2, 6, 12, False), (1, 4, 100, True)])
@pytest.mark.parametrize("clamp,noise_std,white,last", [("relu", 0.0, True, False), ("softplus", 0.5, False, True), ("relu", 0.3, True, True)])
def test_merge_composite_bwd_vs_autograd(ops, B, img, S, two, clamp, noise_std, white, last):
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
    g = torch.Generator().manual_seed(S * 7 + img)
    R = img * img
    coarse = torch.randn((B, R, S, 4), generator=g)
    fine = torch.randn((B, R, S, 4), generator=g)
    for x in (fine, coarse):
        x[..., :3] = torch.sigmoid(x[..., :3])
        x[..., 3] *= 4
    t_c = torch.sort(torch.rand((B, R, S, 1), generator=g) * 1.7 + 0.25, dim=2).values
    t_f = torch.sort(torch.rand((B, R, S, 1), generator=g) * 1.7 + 0.25, dim=2).values
    n = 2 * S if two else S
    noise = torch.randn((B, R, n, 1), generator=g)
    d_pix = torch.randn((B, 3, img, img), generator=g)
    d_dep = torch.randn((B, img, img), generator=g)
    rays, _ = camera_tables((img, img), S, FOV, 0.25, 1.95, "cuda")
    # oracle: autograd through merge + composite + image formatting
    c_r, f_r = coarse.clone().requires_grad_(True), fine.clone().requires_grad_(True)
    if two:
        all_out, all_t, _ = oracle.merge_by_depth(f_r, c_r, t_f, t_c)
    else:
        all_out, all_t = c_r, t_c
    rgb, dist, _ = oracle.composite(all_out, all_t, noise, noise_std, clamp, white, last)
    pixels = rgb.reshape(B, img, img, 3).permute(0, 3, 1, 2) * 2 - 1
    depth = (rays.cpu()[None, :, 2:] * dist).reshape(B, img, img)
    ((pixels * d_pix).sum() + (depth * d_dep).sum()).backward()
    d_fine, d_coarse = ops.merge_composite_bwd(dev(fine) if two else None, dev(coarse), dev(t_f) if two else None, dev(t_c),
                                               dev(noise), rays, dev(d_pix), dev(d_dep), B, img, img, noise_std, clamp, white, last)
    scale = float(c_r.grad.abs().max())
    assert torch.allclose(d_coarse.cpu().view_as(coarse), c_r.grad, rtol=2e-4, atol=2e-5 * max(scale, 1.0)), \
        f"coarse grad: max diff {(d_coarse.cpu().view_as(coarse) - c_r.grad).abs().max().item():.3e} of {scale:.3e}"
    if two:
        assert torch.allclose(d_fine.cpu().view_as(fine), f_r.grad, rtol=2e-4, atol=2e-5 * max(scale, 1.0))
    # only d_pixels / only d_depth
    d_f2, d_c2 = ops.merge_composite_bwd(dev(fine) if two else None, dev(coarse), dev(t_f) if two else None, dev(t_c), dev(noise),
                                         rays, dev(d_pix), None, B, img, img, noise_std, clamp, white, last)
    d_f3, d_c3 = ops.merge_composite_bwd(dev(fine) if two else None, dev(coarse), dev(t_f) if two else None, dev(t_c), dev(noise),
                                         rays, None, dev(d_dep), B, img, img, noise_std, clamp, white, last)
    assert torch.allclose(d_c2 + d_c3, d_coarse, rtol=1e-4, atol=1e-5 * max(scale, 1.0))


@pytest.mark.parametrize("C,D,H,W,N", [(32, 16, 16, 16, 5000), (32, 9, 10, 11, 1000), (8, 5, 6, 7, 300)])
def test_scatter_points_vs_grid_sample_backward(ops, C, D, H, W, N):
    g = torch.Generator().manual_seed(N)
    B = 2
    vol = torch.randn((B, C, D, H, W), generator=g).requires_grad_(True)
    pts = (torch.rand((B, N, 3), generator=g) * 2 - 1) * 0.7
    pts[:, :4] = torch.tensor([[-0.6, -0.6, -0.6], [0.6, 0.6, 0.6], [3, 0, -3], [0, 0, 0]])
    d_feat = torch.randn((B, N, C), generator=g)
    ref = torch.nn.functional.grid_sample(vol, (pts / 0.6).reshape(B, 1, 1, N, 3), mode="bilinear", align_corners=False, padding_mode="border")
    (ref.reshape(B, C, N).permute(0, 2, 1) * d_feat).sum().backward()
    dvol_cl = torch.zeros((B, D, H, W, C), device="cuda")
    ops.scatter_points(dvol_cl, dev(pts), dev(d_feat))
    dvol = ops.volume_from_channels_last(dvol_cl)
    assert torch.allclose(dvol.cpu(), vol.grad, rtol=1e-4, atol=1e-4), (dvol.cpu() - vol.grad).abs().max().item()
    assert torch.equal(ops.volume_to_channels_last(dvol), dvol_cl)


def test_film_sin_elementwise_halves(ops):
    g = torch.Generator().manual_seed(0)
    P, H = 1000, 256
    z = torch.randn((P, H), generator=g)
    bias, freq, phase = torch.randn(H, generator=g) * 0.1, torch.randn(H, generator=g) * 5 + 30, torch.randn(H, generator=g)
    dy = (torch.randn((P, H), generator=g)).to(torch.bfloat16)
    y = ops.film_sin_apply(dev(z), dev(bias), dev(freq), dev(phase))
    u = freq * (z + bias) + phase
    assert y.dtype == torch.bfloat16
    assert (y.float().cpu() - torch.sin(u)).abs().max().item() < 5e-3            # bf16 rounding of a value in [-1, 1]
    dfreq, dphase = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    dz = ops.film_sin_grad(dev(dy), dev(z), dev(bias), dev(freq), dev(phase), dfreq, dphase)
    du = dy.float() * torch.cos(u)
    assert rel_l2(dz.float().cpu(), du * freq) < 4e-3
    assert rel_l2(dphase.cpu(), du.sum(0)) < 1e-4
    assert rel_l2(dfreq.cpu(), (du * (z + bias)).sum(0)) < 1e-4
    # accumulation semantics
    ops.film_sin_grad(dev(dy), dev(z), dev(bias), dev(freq), dev(phase), dfreq, dphase)
    assert rel_l2(dphase.cpu(), 2 * du.sum(0)) < 1e-4


def _oracle_grads(state, siren_type, z, cam, draws, meta, d_pix, d_dep):
    st = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    film = isinstance(z, tuple)
    vol = (z[0] if film else z).clone().requires_grad_(True)
    glob = z[1].clone().requires_grad_(True) if film else None
    out = oracle.render_with_grad(st, siren_type, (vol, glob) if film else vol, cam, draws, **meta)
    ((out["pixels"] * d_pix).sum() + (out["depth"] * d_dep).sum()).backward()
    grads = {k: v.grad for k, v in st.items()}
    grads["volume"] = vol.grad
    if film:
        grads["global"] = glob.grad
    return out, grads


@pytest.mark.parametrize("name", ["fwd_TALLSIREN_FG", "fwd_SHORTSIREN_FG", "fwd_DOUBLESIREN_FG", "fwd_SingleSIREN_dg", "fwd_SHORTSIREN_F",
                                  "fwd_TALLSIREN_dRes", "fwd_TALLSIREN_dResLong", "fwd_SHORTSIREN_FRes"])
def test_generator_backward_vs_oracle_autograd(name):
    """loss = <pixels, G1> + <depth, G2>; gradients w.r.t. every SIREN parameter, the volume and the global feature."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs(name)
    B, img = cam.shape[0], meta["img_size"]
    g = torch.Generator().manual_seed(1)
    d_pix, d_dep = torch.randn((B, 3, img, img), generator=g), torch.randn((B, img, img), generator=g)
    ref_out, ref = _oracle_grads(state, siren_type, z, cam, draws, meta, d_pix, d_dep)
    gen = ImplicitGenerator3d(siren_type, 32 if siren_type in ("TALLSIREN_dRes", "TALLSIREN_dResLong") else 256, 32, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    gen.set_device(torch.device("cuda"))
    gen.siren.precision = "fp32"
    film = isinstance(z, tuple)
    vol = dev(z[0] if film else z).requires_grad_(True)
    glob = dev(z[1]).requires_grad_(True) if film else None
    pixels, depth = gen((vol, glob) if film else vol, dev(cam), draws={k: dev(v) for k, v in draws.items()}, **meta)
    assert pixels.requires_grad and depth.requires_grad
    assert torch.allclose(pixels.detach().cpu(), ref_out["pixels"], atol=2e-3)
    ((pixels * dev(d_pix)).sum() + (depth * dev(d_dep)).sum()).backward()
    torch.cuda.synchronize()
    got = {"siren." + k: p.grad for k, p in gen.siren.named_parameters()}
    got["volume"] = vol.grad
    if film:
        got["global"] = glob.grad
    worst = 0.0
    for k, r in ref.items():
        assert got[k] is not None, f"no gradient for {k}"
        assert got[k].shape == r.shape
        e, c = rel_l2(got[k].cpu(), r), cosine(got[k].cpu(), r)
        worst = max(worst, e)
        print(f"  {name} {k}: rel-L2 {e:.3e} cos {c:.6f} |ref| {float(r.norm()):.3e}")
        assert c > 0.9995 and e < 2e-2, (k, e, c)
    print(f"{name}: worst relative L2 gradient error {worst:.3e}")


def test_siren_boundary_backward_and_amp():
    """gen.siren(points, z, ...) with grad, under autocast + GradScaler-style scaling (utils.py:645-711)."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    state, siren_type, z, cam, draws, meta, taps = fixture_inputs("fwd_DOUBLESIREN_FG")
    gen = ImplicitGenerator3d(siren_type, 256, 32, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    B, S = cam.shape[0], meta["num_steps"]
    pts = dev(taps["points_coarse"].reshape(B, -1, 3))
    vol, glob = dev(z[0]).requires_grad_(True), dev(z[1]).requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        out = gen.siren(pts, (vol.half(), glob), meta["img_size"], S)
        loss = out.float().pow(2).mean() * 1024.0
    loss.backward()
    assert vol.grad is not None and torch.isfinite(vol.grad).all() and float(vol.grad.abs().max()) > 0
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in gen.parameters())
    torch.nn.utils.clip_grad_norm_(gen.parameters(), 1.0)


def test_backward_with_channels_last_3d_volume():
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs("fwd_DOUBLESIREN_FG")
    gen = ImplicitGenerator3d(siren_type, 256, 32, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    gen.siren.precision = "fp32"
    d = {k: dev(v) for k, v in draws.items()}
    grads = []
    for fmt in (torch.contiguous_format, torch.channels_last_3d):
        vol = dev(z[0]).contiguous(memory_format=fmt).requires_grad_(True)
        pixels, depth = gen((vol, dev(z[1])), dev(cam), draws=d, **meta)
        (pixels.sum() + depth.sum()).backward()
        grads.append(vol.grad.contiguous())
    assert torch.allclose(grads[0], grads[1], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("S,clamp,noise_std,white,last", [(12, "relu", 0.0, True, False), (48, "softplus", 0.4, False, True), (100, "relu", 0.2, True, True)])
def test_fancy_integration_autograd(S, clamp, noise_std, white, last):
    """volumetric_rendering.fancy_integration with gradients (cng_composite_bwd) against autograd through the oracle."""
    from conditioned_nerf_gan_b200.generators import volumetric_rendering as vr
    g = torch.Generator().manual_seed(S)
    B, R = 2, 37
    rs = torch.randn((B, R, S, 4), generator=g)
    rs[..., :3] = torch.sigmoid(rs[..., :3])
    rs[..., 3] *= 3
    t = torch.sort(torch.rand((B, R, S, 1), generator=g) * 1.7 + 0.25, dim=2).values
    noise = torch.randn((B, R, S, 1), generator=g)
    g_rgb, g_dist = torch.randn((B, R, 3), generator=g), torch.randn((B, R, 1), generator=g)
    r_ref = rs.clone().requires_grad_(True)
    rgb, dist, _ = oracle.composite(r_ref, t, noise, noise_std, clamp, white, last)
    ((rgb * g_rgb).sum() + (dist * g_dist).sum()).backward()
    r_dev = dev(rs).requires_grad_(True)
    rgb_d, dist_d, w_d = vr.fancy_integration(r_dev, dev(t), "cuda", noise_std=noise_std, last_back=last, white_back=white, clamp_mode=clamp,
                                              noise=dev(noise))
    assert not w_d.requires_grad
    ((rgb_d * dev(g_rgb)).sum() + (dist_d * dev(g_dist)).sum()).backward()
    scale = float(r_ref.grad.abs().max())
    assert torch.allclose(r_dev.grad.cpu(), r_ref.grad, rtol=2e-4, atol=2e-5 * max(scale, 1.0))
