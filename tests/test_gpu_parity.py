"""Parity of every kernel of libcng_b200 (called through the C ABI via conditioned_nerf_gan_b200.ops)
against the torch-CPU oracle on identical inputs and identical random draws.  B200 only (-m gpu).

Tolerances (BASELINE.json north_star): sample_pdf indices, merge order and voxel corner indices
bit-exact; fp32 compositing <= 1e-5 relative; bf16 MLP <= 1e-2 max-abs with PSNR >= 40 dB.
"""
import numpy as np
import pytest
import torch

from conftest import DENSE_FIXTURES, FORWARD_FIXTURES, LIBRARY_FIXTURES, fixture_inputs, latent_fixture_inputs, library_fixture_inputs, load_golden
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


def dev_z(z):
    """z = (volume, global) for the FiLM variants, the volume alone for the unmodulated ones."""
    return tuple(dev(t) for t in z) if isinstance(z, tuple) else dev(z)


# ------------------------------------------------------------------------------------------------
# a4: layout + trilinear lookup
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("shape", [(1, 32, 8, 8, 8), (2, 32, 16, 12, 10), (1, 4, 5, 6, 7), (3, 64, 7, 9, 33)])
def test_channels_last_exact(ops, shape):
    g = torch.Generator().manual_seed(0)
    v = torch.randn(shape, generator=g)
    out = ops.volume_to_channels_last(dev(v)).cpu()
    assert torch.equal(out, v.permute(0, 2, 3, 4, 1).contiguous())


def test_gather_points_golden_and_index_bit_exact(ops):
    fx, _ = load_golden("functions")
    vol, pts = fx["tri/volume"], fx["tri/points"]
    feat, idx = ops.gather_points(ops.volume_to_channels_last(dev(vol)), dev(pts), want_index=True)
    assert torch.allclose(feat.cpu(), fx["tri/features"], rtol=0, atol=2e-6)
    for b in range(vol.shape[0]):
        _, idx_ref = oracle.trilinear_manual(vol[b].numpy(), pts[b].numpy())
        assert np.array_equal(idx[b].cpu().numpy(), idx_ref), "voxel corner index differs from the reference formula"


@pytest.mark.parametrize("C,D,H,W,N", [(32, 64, 64, 64, 20000), (32, 16, 20, 24, 5000), (8, 5, 6, 7, 1000), (32, 32, 32, 32, 1)])
def test_gather_points_random_inside_and_outside(ops, C, D, H, W, N):
    g = torch.Generator().manual_seed(1)
    B = 2
    vol = torch.randn((B, C, D, H, W), generator=g)
    pts = (torch.rand((B, N, 3), generator=g) * 2 - 1) * 0.75       # part of them outside [-0.6, 0.6]^3: border clamp
    pts[:, :8] = torch.tensor([[-0.6, -0.6, -0.6], [0.6, 0.6, 0.6], [0, 0, 0], [0.6, -0.6, 0.0], [5, 5, 5], [-5, 0, 5],
                               [0.6 - 1e-7, 0.59999, -0.59999], [1e-9, -1e-9, 0.3]])[: min(8, N)]
    feat, idx = ops.gather_points(ops.volume_to_channels_last(dev(vol)), dev(pts), want_index=True)
    grid = (pts / 0.6).reshape(B, 1, 1, N, 3)
    ref = torch.nn.functional.grid_sample(vol, grid, mode="bilinear", align_corners=False, padding_mode="border")
    ref = ref.reshape(B, C, N).permute(0, 2, 1)
    assert torch.allclose(feat.cpu(), ref, rtol=0, atol=5e-6)
    _, idx_ref = oracle.trilinear_manual(vol[1].numpy(), pts[1].numpy())
    assert np.array_equal(idx[1].cpu().numpy(), idx_ref)


# ------------------------------------------------------------------------------------------------
# a1-a3 (+a4), a10: fused ray march
# ------------------------------------------------------------------------------------------------
def _ray_setup(B, img, S, V, seed=0):
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
    g = torch.Generator().manual_seed(seed)
    vol = torch.randn((B, 32, V, V, V), generator=g) * 0.3
    cam = oracle.look_at_cam2world(oracle.random_camera_origins(B, 0.7, 1.5, "y", np.random.RandomState(seed)), "y")
    u = torch.rand((B, img * img, S, 1), generator=g)
    rays, t_lin = camera_tables((img, img), S, FOV, 0.25, 1.95, "cuda")
    return vol, cam, u, rays, t_lin


@pytest.mark.parametrize("B,img,S,V", [(2, 16, 6, 16), (1, 64, 12, 32), (3, 10, 5, 8), (2, 32, 24, 64)])
def test_raymarch_coarse_vs_oracle(ops, B, img, S, V):
    vol, cam, u, rays, t_lin = _ray_setup(B, img, S, V)
    feat, t, pts = ops.raymarch_gather_coarse(ops.volume_to_channels_last(dev(vol)), dev(cam), rays, t_lin, dev(u), img, img,
                                              want_points=True)
    p_cam, t_ref, d_cam = oracle.camera_rays(B, S, img, FOV, 0.25, 1.95)
    p_cam, t_ref = oracle.jitter_samples(p_cam, t_ref, d_cam, u)
    p_ref, _, _ = oracle.camera_to_world(p_cam, d_cam, cam)
    assert torch.equal(t.cpu(), t_ref.squeeze(-1)), "jittered distances must be bit-identical (same fp32 ops)"
    assert torch.allclose(pts.cpu(), p_ref, rtol=0, atol=5e-7)      # bmm vs fma chain: <= 2 ulp at |p| <= 2
    # the gather is checked on the kernel's own points, so the 1-ulp position differences do not compound
    f_ref = oracle.trilinear_lookup(vol, pts.cpu().reshape(B, -1, 3), img, S)
    assert torch.allclose(feat.cpu().reshape(B, -1, 32), f_ref, rtol=0, atol=5e-6)


def test_raymarch_coarse_no_jitter_and_odd_image(ops):
    B, img_w, img_h, S, V = 1, 12, 8, 4, 8      # not a multiple of the 8x4 pixel patch in x
    vol, cam, _, _, _ = _ray_setup(B, 8, S, V)
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
    rays, t_lin = camera_tables((img_w, img_h), S, FOV, 0.25, 1.95, "cuda")
    feat, t, pts = ops.raymarch_gather_coarse(ops.volume_to_channels_last(dev(vol)), dev(cam), rays, t_lin, None, img_w, img_h,
                                              want_points=True)
    assert torch.equal(t.cpu(), t_lin.cpu().expand(B, img_w * img_h, S))
    d = rays.cpu()
    p_ref = (d[:, None, :] * t_lin.cpu()[None, :, None]) @ cam[0, :3, :3].T + cam[0, :3, 3]
    assert torch.allclose(pts.cpu()[0], p_ref, rtol=0, atol=1e-6)


@pytest.mark.parametrize("B,img,S,V", [(2, 16, 6, 16), (1, 32, 24, 32)])
def test_raymarch_fine_vs_oracle(ops, B, img, S, V):
    vol, cam, u, rays, _ = _ray_setup(B, img, S, V, seed=3)
    t_fine = 0.25 + 1.7 * u                                                   # [B,R,S,1]
    feat, pts = ops.raymarch_gather_fine(ops.volume_to_channels_last(dev(vol)), dev(cam), rays, dev(t_fine), img, img,
                                         want_points=True)
    _, _, d_cam = oracle.camera_rays(B, S, img, FOV, 0.25, 1.95)
    _, d_w, o_w = oracle.camera_to_world(torch.zeros(B, img * img, S, 3), d_cam, cam)
    p_ref = oracle.fine_points(o_w, d_w, t_fine)
    assert torch.allclose(pts.cpu(), p_ref, rtol=0, atol=5e-7)
    f_ref = oracle.trilinear_lookup(vol, pts.cpu().reshape(B, -1, 3), img, S)
    assert torch.allclose(feat.cpu().reshape(B, -1, 32), f_ref, rtol=0, atol=5e-6)


# ------------------------------------------------------------------------------------------------
# a5-a7: FiLM-SIREN MLP
# ------------------------------------------------------------------------------------------------
def _mlp_setup(siren_type, B, N, feat_std, seed=0):
    state = oracle.init_generator_state(siren_type, seed=seed)
    spec = oracle.SIREN_SPECS[oracle.resolve_siren_type(siren_type)]
    ws = [state[f"siren.{k}.weight"] for k in oracle.layer_keys(siren_type)]
    bs = [state[f"siren.{k}.bias"] for k in oracle.layer_keys(siren_type)]
    g = torch.Generator().manual_seed(seed + 10)
    feat = torch.randn((B, N, 32), generator=g) * feat_std
    glob = torch.randn((B, 256), generator=g) * 0.05 + 0.19
    if spec.get("film", True):
        freq, phase = oracle.film_parameters(glob, state["siren.mapping_network.weight"], state["siren.mapping_network.bias"])
    else:
        freq, phase = torch.ones((B, spec["layers"] * 256)), torch.zeros((B, spec["layers"] * 256))
    fw, fb = state["siren.final_layer.weight"], state["siren.final_layer.bias"]
    ref = oracle.film_siren_mlp(feat, ws, bs, freq, phase, fw, fb, spec["sigmoid_rgb"])
    return spec, ws, bs, feat, freq, phase, fw, fb, ref


def _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb):
    out = ops.film_siren_fwd(dev(feat), [dev(w) for w in ws], [dev(b) for b in bs], dev(freq), dev(phase), dev(fw), dev(fb),
                             spec["sigmoid_rgb"], precision)
    torch.cuda.synchronize()
    return out.cpu()


@pytest.mark.parametrize("siren_type", ["TALLSIREN_FG", "SHORTSIREN_FG", "DOUBLESIREN_FG", "SingleSIREN_dg"])
@pytest.mark.parametrize("B,N", [(2, 1000), (1, 64), (3, 129)])
def test_film_siren_fp32_vs_oracle(ops, siren_type, B, N):
    spec, ws, bs, feat, freq, phase, fw, fb, ref = _mlp_setup(siren_type, B, N, 0.3)
    out = _run_mlp(ops, "fp32", spec, ws, bs, feat, freq, phase, fw, fb)
    err = (out - ref).abs().max().item()
    print(f"{siren_type} fp32 max-abs err {err:.3e}")
    assert err < 5e-4, err       # fp32 accumulation-order differences amplified by freq ~ 30 per layer


# (class, operand format) pairs the host API offers: the frequency_init(12) classes are fp16-only (bf16 operands gave 1.7e-2
# max-abs on SHORTSIREN_FG, over north_star's 1e-2, so that mode was removed; generators/siren.py::_FiLMSirenFG.precision)
TC_CASES = [("TALLSIREN_FG", "bf16"), ("TALLSIREN_FG", "fp16"), ("SHORTSIREN_FG", "fp16"), ("DOUBLESIREN_FG", "bf16"),
            ("DOUBLESIREN_FG", "fp16"), ("SingleSIREN_dg", "bf16"), ("SingleSIREN_dg", "fp16")]


@pytest.mark.parametrize("siren_type,precision", TC_CASES)
@pytest.mark.parametrize("B,N", [(2, 4096), (1, 100), (3, 129), (1, 128 * 300 + 5)])
def test_film_siren_tensor_core_vs_oracle(ops, siren_type, precision, B, N):
    """tcgen05 path, bf16 or fp16 operands.  Features ~ N(0, 0.3^2) (std of a random-init UNet3D output).
    Bar (north_star): 1e-2 max-abs on rgb and sigma, for every offered (class, operand format) pair."""
    spec, ws, bs, feat, freq, phase, fw, fb, ref = _mlp_setup(siren_type, B, N, 0.3)
    out = _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb)
    assert torch.isfinite(out).all()
    err = (out - ref).abs().max().item()
    rms = (out - ref).pow(2).mean().sqrt().item()
    print(f"{siren_type} {precision} B={B} N={N}: max-abs err {err:.3e}, rms {rms:.3e}")
    assert err < 1e-2, err


@pytest.mark.parametrize("siren_type,precision", TC_CASES)
def test_film_siren_tensor_core_stress_features(ops, siren_type, precision):
    """SURVEY.md 8(d) / BASELINE.md 5.5 "stress" case: features ~ N(0, 1) instead of the realistic N(0, 0.3^2).  The layer-0
    pre-activations grow 3.3x; layer 0 runs split (hi/lo) so its own error does not grow, but every later layer sees the
    same sin() outputs, so the error stays at the realistic-case level.  Reported for DESIGN.md; the bar stays 1e-2."""
    spec, ws, bs, feat, freq, phase, fw, fb, ref = _mlp_setup(siren_type, 2, 20000, 1.0, seed=3)
    out = _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb)
    err = (out - ref).abs().max().item()
    rms = (out - ref).pow(2).mean().sqrt().item()
    print(f"STRESS N(0,1) features {siren_type} {precision}: max-abs err {err:.3e}, rms {rms:.3e}")
    assert err < 1e-2, err


def test_film_siren_bf16_matches_fp32_kernel_on_large_batch(ops):
    """Every tile of a multi-wave launch (more tiles than 2 x 148 CTAs) is computed and lands in its slot."""
    spec, ws, bs, feat, freq, phase, fw, fb, _ = _mlp_setup("DOUBLESIREN_FG", 2, 128 * 700 + 17, 0.3)
    a = _run_mlp(ops, "bf16", spec, ws, bs, feat, freq, phase, fw, fb)
    b = _run_mlp(ops, "fp32", spec, ws, bs, feat, freq, phase, fw, fb)
    assert (a - b).abs().max().item() < 1e-2


@pytest.mark.parametrize("siren_type,precision", TC_CASES)
@pytest.mark.parametrize("B,N", [(1, 100), (3, 129), (5, 128 * 9), (2, 128 * 700 + 17)])
def test_film_siren_kernel_organisations_are_bit_identical(ops, siren_type, precision, B, N):
    """The two organisations of the tcgen05 kernel in the default build (film_siren_tc.cu: epilogue warps bound to a tile
    slot / shared between the slots; an experimental build adds the layer-pipelined film_siren_tc3.cu as version 3): same
    operands, same accumulation order, same sine -> the same bits; also within the oracle tolerance.  Covers ragged last
    tiles, item changes inside a CTA's tile sequence (B > 1) and multi-wave launches."""
    import ctypes
    from conditioned_nerf_gan_b200 import _lib
    lib = _lib.load()
    lib.cng_internal_set_tc_version.argtypes = [ctypes.c_int]
    lib.cng_internal_set_tc_version.restype = None
    spec, ws, bs, feat, freq, phase, fw, fb, ref = _mlp_setup(siren_type, B, N, 0.3)
    try:
        lib.cng_internal_set_tc_version(1)
        a = _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb)
        lib.cng_internal_set_tc_version(3)          # default build: falls back to version 1
        b = _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb)
        lib.cng_internal_set_tc_version(2)          # ping-pong kernel with the epilogue warps shared between the slots
        c = _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb)
    finally:
        lib.cng_internal_set_tc_version(0)
    assert torch.equal(a, b), f"max |v1 - v3| = {(a - b).abs().max().item():.3e}"
    assert torch.equal(a, c), f"max |v1 - v2| = {(a - c).abs().max().item():.3e}"
    assert (b - ref).abs().max().item() < 1e-2


@pytest.mark.parametrize("siren_type,precision", TC_CASES)
def test_film_siren_tensor_core_stress_vs_fp32_kernel(ops, siren_type, precision):
    """Race / protocol stress for the mbarrier + named-barrier choreography of the tcgen05 kernel (compute-sanitizer is closed
    on the pool): many seeds x ragged point counts x batch sizes, every launch compared with the exact-fp32 FFMA kernel of
    the same library on the same inputs (<= 1e-2) and with a second launch of itself (bit-identical: a race shows up as a
    run-to-run difference)."""
    worst = 0.0
    for seed, (B, N) in enumerate([(1, 1), (1, 127), (1, 128), (2, 129), (3, 255), (1, 257), (4, 1000), (2, 128 * 149 + 1), (7, 128 * 43 + 77),
                                   (1, 128 * 296), (1, 128 * 297 - 1), (5, 4096), (8, 9999), (2, 128 * 600 + 3)]):
        spec, ws, bs, feat, freq, phase, fw, fb, _ = _mlp_setup(siren_type, B, N, 0.3, seed=100 + seed)
        a = _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb)
        a2 = _run_mlp(ops, precision, spec, ws, bs, feat, freq, phase, fw, fb)
        x = _run_mlp(ops, "fp32", spec, ws, bs, feat, freq, phase, fw, fb)
        assert torch.equal(a, a2), f"run-to-run difference at B={B} N={N}"
        err = (a - x).abs().max().item()
        worst = max(worst, err)
        assert err < 1e-2, (B, N, err)
    print(f"stress {siren_type} {precision}: worst max-abs vs the fp32 kernel over 14 shapes {worst:.3e}")


@pytest.mark.parametrize("precision,tol", [("fp32", 2e-4), ("bf16", 1e-2), ("fp16", 5e-3)])
@pytest.mark.parametrize("B,N", [(1, 100), (3, 129), (2, 128 * 700 + 17)])
def test_film_siren_residual_blocks_vs_oracle(ops, precision, tol, B, N):
    """cng_film_siren_fwd_res: TALLSIREN_dRes as six linear layers with the block input kept and re-added (siren.py:218-230,
    333-408); ragged tiles, several items, more tiles than CTAs (the kept activations live in a per-CTA scratch).  The
    random-init network is nearly constant (outputs ~ 0.02), so the weights are amplified (hidden x6, head x10, features
    ~ N(0,1)) until the residual paths move the output by 0.4 -- a wrong or missing residual cannot hide in the tolerance."""
    spec, ws, bs, feat, freq, phase, fw, fb, _ = _mlp_setup("TALLSIREN_dRes", B, N, 1.0)
    ws = [ws[0] * 12] + [w * 6 for w in ws[1:]]
    fw = fw * 10
    ref = oracle.film_siren_mlp(feat, ws, bs, freq, phase, fw, fb, spec["sigmoid_rgb"], spec["res_save"], spec["res_add"])
    plain = oracle.film_siren_mlp(feat, ws, bs, freq, phase, fw, fb, spec["sigmoid_rgb"])
    assert (ref - plain).abs().max().item() > 0.2
    out = ops.film_siren_fwd(dev(feat), [dev(w) for w in ws], [dev(b) for b in bs], dev(freq), dev(phase), dev(fw), dev(fb),
                             spec["sigmoid_rgb"], precision, spec["res_save"], spec["res_add"]).cpu()
    err = (out - ref).abs().max().item()
    print(f"TALLSIREN_dRes {precision} B={B} N={N}: max-abs err {err:.3e} (residual effect {(ref - plain).abs().max().item():.2f})")
    assert err < tol, err


# ------------------------------------------------------------------------------------------------
# a8: compositing
# ------------------------------------------------------------------------------------------------
def _assert_rel(a, b, rel=1e-5, floor=1e-6, what=""):
    a, b = a.double(), b.double()
    bad = (a - b).abs() > rel * b.abs() + floor
    assert not bad.any(), f"{what}: {int(bad.sum())} elements beyond {rel} relative, worst {(a - b).abs().max().item():.3e}"


def test_composite_golden_cases(ops):
    fx, _ = load_golden("functions")
    for i in range(5):
        cfg = fx[f"comp/case{i}/cfg"]
        rgb, dist, w = ops.composite_fwd(dev(fx["comp/rgb_sigma"]), dev(fx["comp/t"]), dev(fx["comp/noise"]), cfg["noise_std"],
                                         cfg["clamp_mode"], cfg["white_back"], cfg["last_back"])
        _assert_rel(rgb.cpu(), fx[f"comp/case{i}/rgb"], what=f"case{i} rgb")
        _assert_rel(dist.cpu().unsqueeze(-1), fx[f"comp/case{i}/dist"], what=f"case{i} dist")
        _assert_rel(w.cpu().unsqueeze(-1), fx[f"comp/case{i}/weights"], what=f"case{i} weights")


@pytest.mark.parametrize("S", [2, 3, 12, 24, 31, 32, 33, 48, 64, 96, 128, 256, 1000])
@pytest.mark.parametrize("clamp,noise_std,white,last", [("relu", 0.0, True, False), ("softplus", 0.7, False, True)])
def test_composite_vs_oracle(ops, S, clamp, noise_std, white, last):
    g = torch.Generator().manual_seed(S)
    B, R = 2, 301
    rs = torch.randn((B, R, S, 4), generator=g)
    rs[..., :3] = torch.sigmoid(rs[..., :3])
    rs[..., 3] *= 8
    t = torch.sort(torch.rand((B, R, S, 1), generator=g) * 1.7 + 0.25, dim=2).values
    noise = torch.randn((B, R, S, 1), generator=g)
    rgb, dist, w = ops.composite_fwd(dev(rs), dev(t), dev(noise), noise_std, clamp, white, last)
    r_rgb, r_dist, r_w = oracle.composite(rs, t, noise, noise_std, clamp, white, last)
    _assert_rel(rgb.cpu(), r_rgb, what="rgb")
    _assert_rel(dist.cpu(), r_dist.squeeze(-1), what="dist")
    _assert_rel(w.cpu(), r_w.squeeze(-1), what="weights")


def test_composite_properties_full_size(ops):
    """c2-size properties: weights in [0,1], sum <= 1; empty space on a white background is white;
    the result is linear in the colours."""
    n, S = 131072, 48
    g = torch.Generator(device="cuda").manual_seed(0)
    rs = torch.randn((n, S, 4), generator=g, device="cuda")
    rs[..., :3] = torch.sigmoid(rs[..., :3])
    t = torch.sort(torch.rand((n, S), generator=g, device="cuda") * 1.7 + 0.25, dim=1).values
    rgb, dist, w = ops.composite_fwd(rs, t, None, 0.0, "relu", True, False)
    assert (w >= 0).all() and (w <= 1).all() and (w.sum(-1) <= 1 + 1e-5).all()
    assert (dist >= 0).all() and (dist <= t[:, -1] + 1e-5).all()
    rs2 = rs.clone()
    rs2[..., :3] *= 0.5
    rgb2, _, w2 = ops.composite_fwd(rs2, t, None, 0.0, "relu", False, False)
    rgb1, _, _ = ops.composite_fwd(rs, t, None, 0.0, "relu", False, False)
    assert torch.equal(w, w2) and torch.allclose(rgb2, rgb1 * 0.5, rtol=1e-6, atol=1e-7)
    empty = rs.clone()
    empty[..., 3] = -1.0
    rgb0, dist0, w0 = ops.composite_fwd(empty, t, None, 0.0, "relu", True, False)
    assert torch.equal(rgb0, torch.ones_like(rgb0)) and torch.equal(dist0, torch.zeros_like(dist0)) and not w0.any()


# ------------------------------------------------------------------------------------------------
# a9: importance resampling (bit-exact)
# ------------------------------------------------------------------------------------------------
def test_sample_pdf_golden_bit_exact(ops):
    fx, _ = load_golden("functions")
    samples, inds = ops.sample_pdf(dev(fx["pdf/bins"]), dev(fx["pdf/weights"]), dev(fx["pdf/u"]), want_inds=True)
    assert inds.dtype == torch.int64
    assert torch.equal(inds.cpu(), fx["pdf/inds"]), "searchsorted indices differ from the reference's golden vector"
    o_s, o_i, _, _ = oracle.resample_pdf(fx["pdf/bins"], fx["pdf/weights"], fx["pdf/u"])
    assert torch.equal(inds.cpu(), o_i) and torch.equal(samples.cpu(), o_s)


@pytest.mark.parametrize("n,M,K", [(1000, 22, 24), (257, 1, 5), (64, 46, 48), (33, 254, 256), (5, 2047, 64), (3000, 10, 12)])
def test_sample_pdf_vs_oracle_bit_exact(ops, n, M, K):
    g = torch.Generator().manual_seed(M)
    bins = torch.sort(torch.rand((n, M + 1), generator=g) * 1.7 + 0.25, dim=1).values
    w = torch.rand((n, M), generator=g) ** 4
    w[0] = 0                                   # all-zero row: uniform pdf
    w[1] = 0
    w[1, M // 2] = 1                           # spike
    if M > 3:
        w[2, : M // 2] = 0                     # leading empty bins: repeated cdf values
    u = torch.rand((n, K), generator=g)
    u[3, 0], u[3, -1] = 0.0, 1.0 - 2 ** -24    # extremes of torch.rand's range
    samples, inds = ops.sample_pdf(dev(bins), dev(w), dev(u), want_inds=True)
    o_s, o_i, _, _ = oracle.resample_pdf(bins, w, u)
    assert torch.equal(inds.cpu(), o_i), f"{int((inds.cpu() != o_i).sum())} indices differ"
    assert torch.equal(samples.cpu(), o_s), f"max diff {(samples.cpu() - o_s).abs().max().item():.3e}"


@pytest.mark.parametrize("S", [3, 6, 12, 24, 48, 96])
def test_resample_from_coarse_bit_exact(ops, S):
    g = torch.Generator().manual_seed(S)
    n = 2000
    t = torch.sort(torch.rand((n, S), generator=g) * 1.7 + 0.25, dim=1).values
    w = torch.rand((n, S), generator=g) ** 6
    u = torch.rand((n, S), generator=g)
    t_fine, inds = ops.resample_from_coarse(dev(t), dev(w), dev(u), want_inds=True)
    o_s, o_i, _, _ = oracle.coarse_to_fine_t(w.reshape(1, n, S, 1), t.reshape(1, n, S, 1), u, S)
    assert torch.equal(inds.cpu(), o_i) and torch.equal(t_fine.cpu(), o_s)


def test_sample_pdf_properties_full_size(ops):
    """c5-size: 1M rays x 64 bins.  Monotone in u, inside the bin range, deterministic."""
    n, M, K = 1 << 20, 63, 64
    g = torch.Generator(device="cuda").manual_seed(0)
    bins = torch.sort(torch.rand((n, M + 1), generator=g, device="cuda") * 1.7 + 0.25, dim=1).values
    w = torch.rand((n, M), generator=g, device="cuda") ** 4
    u = torch.sort(torch.rand((n, K), generator=g, device="cuda"), dim=1).values
    s1, i1 = ops.sample_pdf(bins, w, u, want_inds=True)
    s2 = ops.sample_pdf(bins, w, u)
    assert torch.equal(s1, s2)
    assert (i1[:, 1:] >= i1[:, :-1]).all() and int(i1.min()) >= 0 and int(i1.max()) <= M + 1
    assert (s1[:, 1:] >= s1[:, :-1] - 1e-6).all()
    assert (s1 >= bins[:, :1] - 1e-6).all() and (s1 <= bins[:, -1:] + 1e-6).all()
    # against the oracle on a slice the CPU finishes in a second
    o_s, o_i, _, _ = oracle.resample_pdf(bins[:4096].cpu(), w[:4096].cpu(), u[:4096].cpu())
    assert torch.equal(i1[:4096].cpu(), o_i) and torch.equal(s1[:4096].cpu(), o_s)


# ------------------------------------------------------------------------------------------------
# a11 + a8 + a12: merge, composite, image formatting
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,img,S", [(2, 8, 6), (1, 16, 24), (2, 4, 48), (1, 4, 100), (1, 4, 128), (1, 3, 200), (1, 3, 256)])
@pytest.mark.parametrize("white,last,noise_std,clamp", [(True, False, 0.0, "relu"), (False, True, 0.5, "softplus")])
def test_merge_composite_vs_oracle(ops, B, img, S, white, last, noise_std, clamp):
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
    g = torch.Generator().manual_seed(S + img)
    R = img * img
    fine = torch.randn((B, R, S, 4), generator=g)
    coarse = torch.randn((B, R, S, 4), generator=g)
    for x in (fine, coarse):
        x[..., :3] = torch.sigmoid(x[..., :3])
        x[..., 3] *= 6
    t_c = torch.sort(torch.rand((B, R, S, 1), generator=g) * 1.7 + 0.25, dim=2).values
    t_f = torch.sort(torch.rand((B, R, S, 1), generator=g) * 1.7 + 0.25, dim=2).values
    t_f[0, 0, :3] = t_c[0, 0, :3]                 # exact ties: stable order puts the fine sample first
    t_f[0, 1] = t_f[0, 1, 0]                      # a run of equal fine distances
    t_c[0, 2] = t_c[0, 2].flip(0)                 # a ray whose coarse distances are NOT sorted: the general sort, not the sort-fine-and-merge path
    noise = torch.randn((B, R, 2 * S, 1), generator=g)
    rays, _ = camera_tables((img, img), S, FOV, 0.25, 1.95, "cuda")
    pixels, depth, taps = ops.merge_composite(dev(fine), dev(coarse), dev(t_f), dev(t_c), dev(noise), rays, B, img, img,
                                              noise_std, clamp, white, last, taps=True)
    all_out, all_t, order = oracle.merge_by_depth(fine, coarse, t_f, t_c)
    assert torch.equal(taps["order"].cpu().long(), order.squeeze(-1)), "merge order differs from the stable sort"
    rgb, dist, _ = oracle.composite(all_out, all_t, noise, noise_std, clamp, white, last)
    _assert_rel(taps["rgb"].cpu(), rgb, what="rgb")
    _assert_rel(taps["dist"].cpu(), dist.squeeze(-1), what="dist")
    pix_ref = rgb.reshape(B, img, img, 3).permute(0, 3, 1, 2) * 2 - 1
    assert torch.allclose(pixels.cpu(), pix_ref, rtol=0, atol=2e-5)
    depth_ref = (rays.cpu()[None, :, 2:] * dist).reshape(B, img, img)
    _assert_rel(depth.cpu(), depth_ref, what="depth")


@pytest.mark.parametrize("n_rays,S", [(100, 24), (7, 100), (33, 256), (1, 1), (50, 12)])
def test_merge_sort_is_torch_stable_sort(ops, n_rays, S):
    """cng_merge_sort: bit-exact order of torch.sort(stable) over cat([fine, coarse]) -- ties (fine == coarse, runs of equal
    fine distances), an unsorted coarse ray, sorted and unsorted fine lists."""
    g = torch.Generator().manual_seed(S)
    t_c = torch.sort(torch.rand((n_rays, S), generator=g) * 1.7 + 0.25, dim=1).values
    t_f = torch.rand((n_rays, S), generator=g) * 1.7 + 0.25
    t_f[0, : min(3, S)] = t_c[0, : min(3, S)]
    if n_rays > 2:
        t_f[1] = t_f[1, 0]
        t_c[2] = t_c[2].flip(0)
    order, t_sorted = ops.merge_sort(dev(t_f), dev(t_c), want_sorted=True)
    cat = torch.cat([t_f, t_c], dim=1)
    ref_t, ref_i = torch.sort(cat, dim=1, stable=True)
    assert torch.equal(order.cpu().long(), ref_i) and torch.equal(t_sorted.cpu(), ref_t)


@pytest.mark.parametrize("S", [2, 16, 17, 32, 33, 48, 64, 65, 128, 129, 255, 256])
@pytest.mark.parametrize("spread", ["narrow", "wide", "signed"])
def test_merge_sort_key_widths(ops, S, spread):
    """The register sort of the fine keys runs on 32-bit keys when the ray's distances span few enough floats and on 64-bit keys
    otherwise (csrc/merge_sort.cuh): both, at every per-lane key count, against torch's stable sort -- distances within a few
    percent of 1 (the generator's ray_start / ray_end: always the 32-bit path), over six decades (always the 64-bit path), and of
    both signs with zeros of both signs; with ties between and inside the lists."""
    g = torch.Generator().manual_seed(S * 7 + len(spread))
    n_rays = 37
    if spread == "narrow":
        t_c = torch.rand((n_rays, S), generator=g) * 0.24 + 0.88
        t_f = torch.rand((n_rays, S), generator=g) * 0.24 + 0.88
    elif spread == "wide":
        t_c = torch.exp(torch.rand((n_rays, S), generator=g) * 14 - 7)
        t_f = torch.exp(torch.rand((n_rays, S), generator=g) * 14 - 7)
    else:
        t_c = torch.randn((n_rays, S), generator=g)
        t_f = torch.randn((n_rays, S), generator=g)
        t_f[3, 0] = 0.0
        t_c[3, 0] = -0.0
    t_c = torch.sort(t_c, dim=1).values
    t_f[0, : min(3, S)] = t_c[0, : min(3, S)]                 # fine == coarse: fine first
    t_f[1] = t_f[1, 0]                                        # all fine distances equal: index order
    t_c[2] = t_c[2, 0]                                        # all coarse distances equal
    order, t_sorted = ops.merge_sort(dev(t_f), dev(t_c), want_sorted=True)
    ref_t, ref_i = torch.sort(torch.cat([t_f, t_c], dim=1), dim=1, stable=True)
    if spread == "signed":
        # -0.0 and +0.0 compare equal for torch; the kernel orders the fine list by bit pattern (-0 before +0).  Row 3 holds the
        # only zeros: compare it by value, the rest bit for bit.
        keep = torch.ones(n_rays, dtype=torch.bool)
        keep[3] = False
        assert torch.equal(t_sorted.cpu()[3], ref_t[3])
        assert torch.equal(order.cpu().long()[keep], ref_i[keep]) and torch.equal(t_sorted.cpu()[keep], ref_t[keep])
    else:
        assert torch.equal(order.cpu().long(), ref_i) and torch.equal(t_sorted.cpu(), ref_t)


@pytest.mark.parametrize("S", [24, 64, 100, 200, 256])
def test_merge_composite_unsorted_fine_vs_oracle(ops, S):
    """The merge as the generator calls it: the fine distances arrive in the order of the random draws (unsorted), at every
    chunking of the compositing loop (2S <= 128: one pass; beyond: chunks of 128 with the transmittance carried)."""
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
    g = torch.Generator().manual_seed(S)
    B, img = 1, 5
    R = img * img
    fine = torch.randn((B, R, S, 4), generator=g)
    coarse = torch.randn((B, R, S, 4), generator=g)
    for x in (fine, coarse):
        x[..., :3] = torch.sigmoid(x[..., :3])
        x[..., 3] *= 6
    t_c = torch.sort(torch.rand((B, R, S, 1), generator=g) * 0.24 + 0.88, dim=2).values
    t_f = torch.rand((B, R, S, 1), generator=g) * 0.24 + 0.88
    noise = torch.randn((B, R, 2 * S, 1), generator=g)
    rays, _ = camera_tables((img, img), S, FOV, 0.88, 1.12, "cuda")
    for white, last, noise_std, clamp in ((True, False, 0.0, "relu"), (False, True, 0.5, "softplus")):
        pixels, depth, taps = ops.merge_composite(dev(fine), dev(coarse), dev(t_f), dev(t_c), dev(noise), rays, B, img, img,
                                                  noise_std, clamp, white, last, taps=True)
        all_out, all_t, order = oracle.merge_by_depth(fine, coarse, t_f, t_c)
        assert torch.equal(taps["order"].cpu().long(), order.squeeze(-1)), "merge order differs from the stable sort"
        rgb, dist, _ = oracle.composite(all_out, all_t, noise, noise_std, clamp, white, last)
        _assert_rel(taps["rgb"].cpu(), rgb, what="rgb")
        _assert_rel(taps["dist"].cpu(), dist.squeeze(-1), what="dist")
        assert torch.isfinite(depth).all() and torch.isfinite(pixels).all()


def test_order_and_index_kernels_random_stress(ops):
    """40 random rounds of tools/gpu/stress_c5.py: merge order against torch's stable sort and sample_pdf against the oracle, bit
    for bit, over random sample counts, value ranges spanning many binades, heavy ties, unsorted coarse lists, empty bins."""
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("stress_c5", os.path.join(os.path.dirname(__file__), "..", "tools", "gpu", "stress_c5.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    assert mod.run(40, seed=99) == 0


def test_merge_composite_rejects_more_than_512_samples(ops):
    """Documented limit of the per-warp sort (include/cng_b200.h): 2S <= 512 samples per ray; beyond it the call fails loudly."""
    from conditioned_nerf_gan_b200._lib import CngError
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
    S, img = 300, 2
    x = torch.rand((1, img * img, S, 4), device="cuda")
    t = torch.rand((1, img * img, S, 1), device="cuda").sort(dim=2).values
    rays, _ = camera_tables((img, img), S, FOV, 0.25, 1.95, "cuda")
    with pytest.raises(CngError, match="512"):
        ops.merge_composite(x, x, t, t, None, rays, 1, img, img, 0.0, "relu", True, False)


def test_merge_composite_coarse_only(ops):
    from conditioned_nerf_gan_b200.generators.volumetric_rendering import camera_tables
    g = torch.Generator().manual_seed(4)
    B, img, S = 2, 8, 12
    coarse = torch.randn((B, img * img, S, 4), generator=g)
    t_c = torch.sort(torch.rand((B, img * img, S, 1), generator=g) + 0.25, dim=2).values
    rays, _ = camera_tables((img, img), S, FOV, 0.25, 1.95, "cuda")
    pixels, depth, taps = ops.merge_composite(None, dev(coarse), None, dev(t_c), None, rays, B, img, img, 0.0, "relu", True, False, taps=True)
    rgb, dist, _ = oracle.composite(coarse, t_c, torch.zeros_like(t_c), 0.0, "relu", True, False)
    _assert_rel(taps["rgb"].cpu(), rgb, what="rgb")
    assert torch.equal(taps["order"].cpu().long(), torch.arange(S).expand(B, img * img, S))


# ------------------------------------------------------------------------------------------------
# whole forward through ImplicitGenerator3d against the golden vectors recorded from the reference
# ------------------------------------------------------------------------------------------------
def _generator(siren_type, state, precision):
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    z_dim = 32 if siren_type in ("TALLSIREN_dRes", "TALLSIREN_dResLong") else 256      # those classes read z_dim features (configs/thousand/direct_volume/dRes.py)
    gen = ImplicitGenerator3d(siren_type, z_dim, 32, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    gen.set_device(torch.device("cuda"))
    gen.siren.precision = precision
    return gen


FP16_ONLY = ("SHORTSIREN_FG", "SHORTSIREN_dg", "SHORTSIREN_F", "SHORTSIREN_FRes")       # classes whose tensor-core mode is fp16 operands


def _fixture_cases(names):
    out = []
    for n in names:
        for precision in ("fp32", "bf16", "fp16"):
            if precision == "bf16" and any(n.endswith(c) for c in FP16_ONLY):
                continue
            out.append((n, precision))
    return out


def _render_fixture(name, precision):
    state, siren_type, z, cam, draws, meta, taps = fixture_inputs(name)
    gen = _generator(siren_type, state, precision)
    zc = dev_z(z)
    vol_d, glob_d = gen.siren.split_z(zc)
    d = {k: dev(v) for k, v in draws.items()}
    with torch.no_grad():
        out = gen._render(vol_d, glob_d, dev(cam), meta["img_size"], meta["fov"], meta["ray_start"], meta["ray_end"],
                          meta["num_steps"], meta["hierarchical_sample"], dict(meta, draws=d), taps=True)
        pixels, depth = gen(zc, dev(cam), draws=d, **meta)
    torch.cuda.synchronize()
    assert torch.equal(pixels, out["pixels"]) and torch.equal(depth, out["depth"]), "forward is not deterministic"
    B, img, S = cam.shape[0], meta["img_size"], meta["num_steps"]
    assert pixels.shape == (B, 3, img, img) and depth.shape == (B, img, img) and pixels.is_contiguous()
    assert torch.allclose(out["points_coarse"].cpu(), taps["points_coarse"].reshape(B, -1, S, 3), rtol=0, atol=5e-7)
    return state, siren_type, z, cam, draws, meta, taps, out, pixels.cpu(), depth.cpu()


@pytest.mark.parametrize("name,precision", _fixture_cases(FORWARD_FIXTURES))
def test_forward_vs_reference_golden(name, precision):
    """Random-init fixtures recorded from the reference (near-empty scenes: sigma ~ 1e-2, every relu-mode pixel is decided by
    the far-plane sample, oracle.far_plane_sigma).  Asserted: the MLP bar (1e-2 max-abs in tensor-core modes) on the coarse
    pass and the whole-image PSNR / pixel / depth bars in fp32 mode.  The reduced-precision image PSNR of these scenes is
    printed, not asserted: the image-level bar is asserted on the fixtures WITH density (test_forward_vs_reference_dense)."""
    state, siren_type, z, cam, draws, meta, taps, out, pixels, depth = _render_fixture(name, precision)
    B, img, S = cam.shape[0], meta["img_size"], meta["num_steps"]
    mlp_tol = 5e-4 if precision == "fp32" else 1e-2
    err_c = (out["rgb_sigma_coarse"].cpu() - taps["rgb_sigma_coarse"].reshape(B, -1, S, 4)).abs().max().item()
    err_p = (pixels - taps["pixels"]).abs().max().item()
    psnr_full = oracle.psnr(pixels, taps["pixels"])
    print(f"{name} {precision}: coarse rgb_sigma max-abs {err_c:.3e}; pixels max-abs {err_p:.3e}, PSNR {psnr_full:.1f} dB (whole image)")
    assert err_c < mlp_tol
    if precision == "fp32":
        assert psnr_full >= 60.0 and err_p < 2e-3
        assert torch.allclose(depth, taps["depth"], rtol=0, atol=2e-3)
    elif meta["clamp_mode"] == "softplus":
        assert psnr_full >= 40.0            # continuous clamp: the whole-image bar holds on the near-empty scenes too


@pytest.mark.parametrize("name,precision", _fixture_cases(DENSE_FIXTURES))
def test_forward_vs_reference_dense(name, precision):
    """Fixtures with real density recorded from the reference (head rows of the random-init network scaled,
    oracle.DENSE_HEAD_GAINS: alpha spans 0..~0.9, rays saturate, occlusion decides the pixel).  All clamp_mode "relu".
    Whole image, every pixel: PSNR >= 40 dB in tensor-core modes (north_star), >= 60 dB in fp32 mode.  The MLP bar is
    1e-2 max-abs on the colours and 1e-2 x sigma gain on sigma (the same relative error of the same hidden activations)."""
    state, siren_type, z, cam, draws, meta, taps, out, pixels, depth = _render_fixture(name, precision)
    B, img, S = cam.shape[0], meta["img_size"], meta["num_steps"]
    sigma_gain, rgb_gain = oracle.DENSE_HEAD_GAINS[oracle.resolve_siren_type(siren_type)][:2]
    d_c = (out["rgb_sigma_coarse"].cpu() - taps["rgb_sigma_coarse"].reshape(B, -1, S, 4)).abs()
    err_rgb, err_sigma = d_c[..., :3].max().item(), d_c[..., 3].max().item()
    psnr = oracle.psnr(pixels, taps["pixels"])
    err_p = (pixels - taps["pixels"]).abs().max().item()
    err_d = (depth - taps["depth"]).abs().max().item()
    w = taps["weights_final"][..., 0]
    print(f"{name} {precision}: coarse rgb max-abs {err_rgb:.3e}, sigma {err_sigma:.3e} (gain {sigma_gain:g}); pixels max-abs {err_p:.3e}, "
          f"depth {err_d:.3e}, PSNR {psnr:.1f} dB whole image; reference far-plane weight mean {float(w[..., -1].mean()):.4f}")
    if precision == "fp32":
        assert err_rgb < 5e-4 * max(1.0, rgb_gain) and err_sigma < 5e-4 * sigma_gain
        assert psnr >= 60.0
        assert err_d < 5e-3
    else:
        assert err_rgb < 1e-2 * (rgb_gain if not oracle.SIREN_SPECS[oracle.resolve_siren_type(siren_type)]["sigmoid_rgb"] else 1.0)
        assert err_sigma < 1e-2 * sigma_gain
        assert psnr >= 40.0


@pytest.mark.parametrize("siren_type", ["TALLSIREN_FG", "SHORTSIREN_FG", "DOUBLESIREN_FG"])
def test_bf16_image_psnr_softplus(siren_type):
    """BASELINE north_star: the tensor-core MLP (each class's default operand format: bf16, fp16 for SHORTSIREN_FG) keeps
    PSNR >= 40 dB on rendered images.  32x32, 12+12 samples, 32^3 volume, clamp_mode softplus (continuous everywhere, see
    oracle.far_plane_sigma), whole image, no pixel excluded."""
    B, img, S, V = 2, 32, 12, 32
    state = oracle.init_generator_state(siren_type, seed=5)
    g = torch.Generator().manual_seed(6)
    z = (torch.randn((B, 32, V, V, V), generator=g) * 0.3, torch.randn((B, 256), generator=g) * 0.05 + 0.19)
    cam = oracle.look_at_cam2world(oracle.random_camera_origins(B, 0.7, 1.5, "y", np.random.RandomState(7)), "y")
    draws = oracle.draw_randoms(B, img, S, True, g)
    meta = dict(img_size=img, fov=FOV, ray_start=0.25, ray_end=1.95, num_steps=S, hierarchical_sample=True,
                clamp_mode="softplus", nerf_noise=0.0, white_back=True)
    ref = oracle.render(state, siren_type, z, cam, draws, **meta)
    gen = _generator(siren_type, state, "fp16" if siren_type.startswith("SHORT") else "bf16")
    with torch.no_grad():
        pixels, depth = gen((dev(z[0]), dev(z[1])), dev(cam), draws={k: dev(v) for k, v in draws.items()}, **meta)
    psnr = oracle.psnr(pixels.cpu(), ref["pixels"])
    err = (pixels.cpu() - ref["pixels"]).abs().max().item()
    print(f"{siren_type} bf16 softplus 32x32: PSNR {psnr:.1f} dB, max-abs pixel err {err:.3e}, depth err {(depth.cpu() - ref['depth']).abs().max().item():.3e}")
    assert psnr >= 40.0


def test_forward_draws_follow_reference_rng_order():
    """Without replayed draws the generator consumes torch's CUDA generator in the reference's order and
    shapes (rand[B,R,S,1], randn[B,R,S,1], rand[B*R,S], randn[B,R,2S,1])."""
    state, siren_type, z, cam, _, meta, _ = fixture_inputs("fwd_DOUBLESIREN_FG")
    gen = _generator(siren_type, state, "fp32")
    B, R, S = cam.shape[0], meta["img_size"] ** 2, meta["num_steps"]
    zc = (dev(z[0]), dev(z[1]))
    torch.manual_seed(123)
    with torch.no_grad():
        a, _ = gen(zc, dev(cam), **meta)
    torch.manual_seed(123)
    d = {"u_jitter": torch.rand((B, R, S, 1), device="cuda"), "noise_coarse": torch.randn((B, R, S, 1), device="cuda"),
         "u_resample": torch.rand((B * R, S), device="cuda"), "noise_final": torch.randn((B, R, 2 * S, 1), device="cuda")}
    with torch.no_grad():
        b, _ = gen(zc, dev(cam), draws=d, **meta)
    assert torch.equal(a, b)


def test_siren_secondary_boundary(ops):
    """gen.siren(points, z, img_size, num_steps) as called by extract_shapes.py:63-68."""
    state, siren_type, z, cam, draws, meta, taps = fixture_inputs("fwd_TALLSIREN_FG")
    gen = _generator(siren_type, state, "fp32")
    B, S, img = cam.shape[0], meta["num_steps"], meta["img_size"]
    pts = taps["points_coarse"].reshape(B, -1, 3)
    with torch.no_grad():
        out = gen.siren(dev(pts), (dev(z[0]), dev(z[1])), img, S)
    assert out.shape == (B, pts.shape[1], 4)
    assert (out.cpu() - taps["rgb_sigma_coarse"].reshape(B, -1, 4)).abs().max().item() < 5e-4


def test_staged_forward_matches_forward():
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs("fwd_TALLSIREN_FG")
    gen = _generator(siren_type, state, "fp32")
    meta = dict(meta, nerf_noise=0.0)
    B = cam.shape[0]
    d = {k: dev(v) for k, v in draws.items()}
    zc = (dev(z[0]), dev(z[1]))
    with torch.no_grad():
        full, depth_full = gen(zc, dev(cam), draws=d, **meta)
    R, S = meta["img_size"] ** 2, meta["num_steps"]
    for b in range(B):
        db = {"u_jitter": d["u_jitter"][b:b + 1], "noise_coarse": d["noise_coarse"][b:b + 1],
              "u_resample": d["u_resample"][b * R:(b + 1) * R], "noise_final": d["noise_final"][b:b + 1]}
        px, dp = gen.staged_forward((zc[0][b:b + 1], zc[1][b:b + 1]), dev(cam[b:b + 1]), max_batch_size=1, draws=db, **meta)
        assert torch.equal(px[0], full[b]) and torch.equal(dp[0], depth_full[b])
    # one object, several poses, chunked
    poses = dev(cam[:1]).expand(3, 4, 4).contiguous()
    torch.manual_seed(0)
    px, dp = gen.staged_forward((zc[0][:1], zc[1][:1]), poses, max_batch_size=2, **meta)
    assert px.shape == (3, 3, meta["img_size"], meta["img_size"]) and torch.isfinite(px).all()


def test_errors_on_device(ops):
    with pytest.raises(TypeError):
        ops.composite_fwd(torch.zeros(2, 4, 4, device="cuda"), torch.zeros(2, 4, device="cuda"), None, 0.0, "tanh")
    with pytest.raises(ValueError):
        ops.sample_pdf(torch.zeros(2, 4, device="cuda"), torch.zeros(2, 4, device="cuda"), torch.zeros(2, 3, device="cuda"))
    from conditioned_nerf_gan_b200 import _lib
    v = torch.zeros((1, 4, 4, 4, 6), device="cuda")     # C = 6: not a multiple of 4
    with pytest.raises(_lib.CngError):
        ops.gather_points(v, torch.zeros((1, 3, 3), device="cuda"))
    # empty batch / empty point set: no-op
    out = ops.composite_fwd(torch.zeros(0, 4, 4, device="cuda"), torch.zeros(0, 4, device="cuda"), None, 0.0, "relu")
    assert out[0].shape == (0, 3)


def test_streaming_host_batches_match_direct_calls():
    """render_host_batches (three streams, buffers in flight) returns, in order, exactly what direct calls return."""
    from conditioned_nerf_gan_b200.streaming import render_host_batches
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs("fwd_DOUBLESIREN_FG")
    gen = _generator(siren_type, state, "fp32")
    meta = dict(meta, nerf_noise=0.0)
    d = {k: dev(v) for k, v in draws.items()}
    batches = []
    for k in range(5):
        g = torch.Generator().manual_seed(k)
        batches.append(((z[0] + 0.01 * k).pin_memory(), z[1].pin_memory(), cam.pin_memory()))
    expect = []
    with torch.no_grad():
        for vol, glob, c in batches:
            px, dp = gen((dev(vol), dev(glob)), dev(c), draws=d, **meta)
            expect.append((px.cpu(), dp.cpu()))
    got = [(px.clone(), dp.clone()) for px, dp in render_host_batches(gen, batches, dict(meta, draws=d))]
    assert len(got) == len(expect)
    for (a, b), (c, e) in zip(got, expect):
        assert torch.equal(a, c) and torch.equal(b, e)
    assert list(render_host_batches(gen, [], dict(meta, draws=d))) == []


def test_channels_last_3d_volume_needs_no_layout_kernel(ops):
    """An encoder output in torch.channels_last_3d memory format is consumed in place (no layout kernel, same image)."""
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs("fwd_TALLSIREN_FG")
    gen = _generator(siren_type, state, "fp32")
    d = {k: dev(v) for k, v in draws.items()}
    vol = dev(z[0])
    vol_cl3d = vol.contiguous(memory_format=torch.channels_last_3d)
    assert not vol_cl3d.is_contiguous()
    n0 = ops.launch_count
    with torch.no_grad():
        a, da = gen((vol, dev(z[1])), dev(cam), draws=d, **meta)
    n1 = ops.launch_count
    with torch.no_grad():
        b, db = gen((vol_cl3d, dev(z[1])), dev(cam), draws=d, **meta)
    n2 = ops.launch_count
    assert torch.equal(a, b) and torch.equal(da, db)
    assert (n2 - n1) == (n1 - n0) - 1
    assert ops.volume_to_channels_last(vol_cl3d).data_ptr() == vol_cl3d.data_ptr()


def _full_size_case(B, img, S, V, seed, dense, n_items=None):
    """Seeded inputs of a BASELINE config at full size; ``dense``: TALLSIREN_FG with the dense head (oracle.DENSE_HEAD_GAINS)."""
    siren_type = "TALLSIREN_FG"
    state = oracle.init_generator_state(siren_type, seed=seed)
    if dense:
        state = oracle.dense_head_state(state, *oracle.DENSE_HEAD_GAINS[siren_type])
    g = torch.Generator().manual_seed(seed + 1)
    z = (torch.randn((B, 32, V, V, V), generator=g) * 0.3, torch.randn((B, 256), generator=g) * 0.05 + 0.19)
    cam = oracle.look_at_cam2world(oracle.random_camera_origins(B, 0.7, 1.5, "y", np.random.RandomState(seed + 2)), "y")
    draws = oracle.draw_randoms(B, img, S, True, g)
    meta = dict(img_size=img, fov=FOV, ray_start=0.25, ray_end=1.95, num_steps=S, hierarchical_sample=True,
                clamp_mode="relu", nerf_noise=0.0, white_back=True)
    return siren_type, state, z, cam, draws, meta


def _slice_item(z, cam, draws, i, R):
    return ((z[0][i:i + 1], z[1][i:i + 1]), cam[i:i + 1],
            {"u_jitter": draws["u_jitter"][i:i + 1], "noise_coarse": draws["noise_coarse"][i:i + 1],
             "u_resample": draws["u_resample"][i * R:(i + 1) * R], "noise_final": draws["noise_final"][i:i + 1]})


def _check_indexing_bit_exact(out, draws_i, S, item):
    """sample_pdf indices / samples and the merge order, bit-exact GIVEN the kernel's own upstream values (its coarse weights
    differ from the oracle's in the last ulps; near-ties may then legitimately resolve differently)."""
    B1 = 1
    t_c = out["t_coarse"][item:item + 1].cpu()
    w_c = out["weights_coarse"][item:item + 1].cpu()
    t_f = out["t_fine"][item:item + 1].cpu().reshape(B1, -1, S)
    o_s, o_i, _, _ = oracle.coarse_to_fine_t(w_c.unsqueeze(-1), t_c.unsqueeze(-1), draws_i["u_resample"], S)
    R = t_c.shape[1]
    assert torch.equal(out["resample_inds"].cpu().reshape(-1, R, S)[item].reshape(-1, S), o_i), "sample_pdf bin indices"
    assert torch.equal(t_f.reshape(-1, S), o_s), "sample_pdf samples"
    order = torch.sort(torch.cat([t_f, t_c], dim=-1), dim=-1, stable=True).indices
    assert torch.equal(out["merge_order"][item:item + 1].cpu().long(), order), "merge order differs from the stable sort"


@pytest.mark.parametrize("dense", [False, True])
def test_config1_full_size_vs_oracle(dense):
    """BASELINE configs[0] at full size: 64x64, 12+12 samples, batch 1, 32^3 x 32 volume, TALLSIREN_FG -- the reference's own
    CPU-runnable case, rendered by the CUDA path (fp32 and bf16 modes) and by the oracle on the same draws.  ``dense``: the
    same network with density (whole-image PSNR asserted in every mode); random init: the MLP bar in every mode, the image
    bars in fp32 mode (its relu-mode pixels are all far-plane decisions, see oracle.far_plane_sigma)."""
    B, img, S, V = 1, 64, 12, 32
    siren_type, state, z, cam, draws, meta = _full_size_case(B, img, S, V, 21, dense)
    ref = oracle.render(state, siren_type, z, cam, draws, **meta)
    d = {k: dev(v) for k, v in draws.items()}
    sigma_gain = oracle.DENSE_HEAD_GAINS[siren_type][0] if dense else 1.0
    for precision, min_psnr in (("fp32", 60.0), ("bf16", 40.0), ("fp16", 40.0)):
        gen = _generator(siren_type, state, precision)
        with torch.no_grad():
            out = gen._render(dev(z[0]), dev(z[1]), dev(cam), img, FOV, 0.25, 1.95, S, True, dict(meta, draws=d), taps=True)
        pixels = out["pixels"].cpu()
        assert torch.equal(out["t_coarse"].cpu(), ref["t_coarse"].squeeze(-1))
        _check_indexing_bit_exact(out, draws, S, 0)
        d_c = (out["rgb_sigma_coarse"].cpu() - ref["rgb_sigma_coarse"]).abs()
        tol = 5e-4 if precision == "fp32" else 1e-2
        assert d_c[..., :3].max().item() < tol and d_c[..., 3].max().item() < tol * sigma_gain
        psnr = oracle.psnr(pixels, ref["pixels"])
        print(f"config 1 {'dense' if dense else 'random-init'} {precision}: PSNR {psnr:.1f} dB whole image, coarse rgb err {d_c[..., :3].max().item():.2e}, "
              f"sigma err {d_c[..., 3].max().item():.2e}")
        if precision == "fp32":
            assert (out["merge_order"].cpu().long() != ref["merge_order"].squeeze(-1)).float().mean().item() < 2e-3
            assert psnr >= min_psnr
            assert torch.allclose(out["depth"].cpu(), ref["depth"], atol=5e-3)
        elif dense:
            assert psnr >= min_psnr


def test_config2_full_size_vs_oracle():
    """BASELINE configs[1] -- the headline workload -- at FULL size against the oracle: batch 8, 128x128, 24+24 samples, 64^3
    volume, TALLSIREN_FG with density.  The CUDA path renders the whole batch (bf16 tensor-core mode, the benchmarked one, and
    fp32 mode); the oracle renders images 2 and 6 of the 8 on the same draws (~2 s each).  Asserted per image: coarse MLP
    <= 1e-2 (sigma: x gain), sample_pdf indices / samples and merge order bit-exact given the kernel's own upstream values,
    PSNR >= 40 dB over the whole image (fp32 mode: >= 60 dB), depth; and the batch render equals the single-image render
    bit for bit (rays and images are independent: the sharding argument of DESIGN.md 6)."""
    B, img, S, V = 8, 128, 24, 64
    R = img * img
    siren_type, state, z, cam, draws, meta = _full_size_case(B, img, S, V, 40, True)
    sigma_gain = oracle.DENSE_HEAD_GAINS[siren_type][0]
    d = {k: dev(v) for k, v in draws.items()}
    zc, camc = (dev(z[0]), dev(z[1])), dev(cam)
    refs = {}
    for i in (2, 6):
        zi, ci, di = _slice_item(z, cam, draws, i, R)
        refs[i] = (oracle.render(state, siren_type, zi, ci, di, **meta), di)
    for precision, min_psnr in (("bf16", 40.0), ("fp32", 60.0)):
        gen = _generator(siren_type, state, precision)
        with torch.no_grad():
            out = gen._render(zc[0], zc[1], camc, img, FOV, 0.25, 1.95, S, True, dict(meta, draws=d), taps=True)
            a, da = gen(zc, camc, draws=d, **meta)
        assert torch.equal(a, out["pixels"]) and torch.isfinite(a).all() and torch.isfinite(da).all()
        assert float(a.min()) >= -1 - 1e-5 and float(a.max()) <= 1 + 1e-5
        for i, (ref, di) in refs.items():
            assert torch.equal(out["t_coarse"][i].cpu(), ref["t_coarse"][0].squeeze(-1))
            _check_indexing_bit_exact(out, di, S, i)
            d_c = (out["rgb_sigma_coarse"][i].cpu() - ref["rgb_sigma_coarse"][0]).abs()
            tol = 5e-4 if precision == "fp32" else 1e-2
            psnr = oracle.psnr(a[i].cpu(), ref["pixels"][0])
            err_d = (da[i].cpu() - ref["depth"][0]).abs()
            w_last = float(ref["weights_final"][..., -1, 0].mean())
            print(f"config 2 image {i} {precision}: PSNR {psnr:.1f} dB whole image, coarse rgb err {d_c[..., :3].max().item():.2e}, sigma err "
                  f"{d_c[..., 3].max().item():.2e} (gain {sigma_gain:g}), depth err max {err_d.max().item():.2e} mean {err_d.mean().item():.2e}, "
                  f"reference far-plane weight mean {w_last:.4f}")
            assert d_c[..., :3].max().item() < tol and d_c[..., 3].max().item() < tol * sigma_gain
            assert psnr >= min_psnr
            if precision == "fp32":
                assert err_d.mean().item() < 1e-4
            # the same image rendered alone: bit-identical
            zi, ci, dii = _slice_item(zc, camc, d, i, R)
            with torch.no_grad():
                s1, ds1 = gen(zi, ci, draws=dii, **meta)
            assert torch.equal(s1[0], a[i]) and torch.equal(ds1[0], da[i])


@pytest.mark.parametrize("precision", ["bf16", "fp16"])
def test_fused_gather_forward_is_bit_identical(ops, precision):
    """cng_film_siren_fwd_gather (trilinear lookup fused into K2's prologue, no feat[B,N,32] in HBM) against the separate
    K1 gather + K2: the same arithmetic in the same order, hence the same image bit for bit; ragged tiles and a shared volume."""
    import ctypes
    from conditioned_nerf_gan_b200 import _lib
    lib = _lib.load()
    lib.cng_internal_set_fused_gather.argtypes = [ctypes.c_int]
    lib.cng_internal_set_fused_gather.restype = None
    for (B, img, S, V) in ((2, 20, 7, 16), (1, 64, 12, 32)):
        siren_type, state, z, cam, draws, meta = _full_size_case(B, img, S, V, 70 + B, True)
        gen = _generator(siren_type, state, precision)
        d = {k: dev(v) for k, v in draws.items()}
        outs = []
        try:
            for mode in (0, 1):
                lib.cng_internal_set_fused_gather(mode)
                with torch.no_grad():
                    outs.append(gen((dev(z[0]), dev(z[1])), dev(cam), draws=d, **meta))
        finally:
            lib.cng_internal_set_fused_gather(-1)
        assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1]), (outs[0][0] - outs[1][0]).abs().max().item()
    # the entry point itself against gather_points + film_siren_fwd
    spec, ws, bs, feat, freq, phase, fw, fb, _ = _mlp_setup("TALLSIREN_FG", 2, 777, 0.3)
    g = torch.Generator().manual_seed(5)
    vol = dev(torch.randn((2, 32, 9, 10, 11), generator=g))
    pts = dev((torch.rand((2, 777, 3), generator=g) - 0.5) * 1.6)
    vol_cl = ops.volume_to_channels_last(vol)
    ref = ops.film_siren_fwd(ops.gather_points(vol_cl, pts), [dev(w) for w in ws], [dev(b) for b in bs], dev(freq), dev(phase), dev(fw), dev(fb),
                             spec["sigmoid_rgb"], precision)
    out = ops.film_siren_fwd_gather(vol_cl, pts, [dev(w) for w in ws], [dev(b) for b in bs], dev(freq), dev(phase), dev(fw), dev(fb),
                                    spec["sigmoid_rgb"], precision)
    assert torch.equal(out, ref)


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp16"])
def test_latent_shortsiren_forward_vs_reference(precision):
    """SHORTSIREN (siren.py:1172-1224, the generator of the default config): sample positions as the MLP input, FiLM parameters from a
    latent vector through CustomMappingNetwork.  K1 runs in points-only mode, the positions ride in 3 of the 32 operand channels."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    state, latent, cam, draws, meta, taps = latent_fixture_inputs()
    gen = ImplicitGenerator3d("SHORTSIREN", latent.shape[1], 3, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    gen.set_device(torch.device("cuda"))
    gen.siren.precision = precision
    d = {k: dev(v) for k, v in draws.items()}
    B, img, S = cam.shape[0], meta["img_size"], meta["num_steps"]
    with torch.no_grad():
        out = gen._render(None, dev(latent), dev(cam), img, meta["fov"], meta["ray_start"], meta["ray_end"], S, True, dict(meta, draws=d), taps=True)
        pixels, depth = gen(dev(latent), dev(cam), draws=d, **meta)
        rs = gen.siren(dev(taps["points_coarse"].reshape(B, -1, 3)), dev(latent), img, S)       # the secondary boundary
    assert torch.equal(pixels, out["pixels"])
    assert torch.allclose(out["points_coarse"].cpu(), taps["points_coarse"], rtol=0, atol=5e-7)
    d_c = (out["rgb_sigma_coarse"].cpu() - taps["rgb_sigma_coarse"]).abs()
    tol = 5e-4 if precision == "fp32" else 1e-2
    psnr = oracle.psnr(pixels.cpu(), taps["pixels"])
    print(f"SHORTSIREN {precision}: coarse rgb err {d_c[..., :3].max().item():.2e}, sigma err {d_c[..., 3].max().item():.2e} (gain 300), PSNR {psnr:.1f} dB")
    assert d_c[..., :3].max().item() < tol and d_c[..., 3].max().item() < tol * 300
    assert (rs.cpu().reshape(B, -1, S, 4) - taps["rgb_sigma_coarse"]).abs()[..., :3].max().item() < tol
    assert psnr >= (60.0 if precision == "fp32" else 40.0)


def test_second_device_renders_like_the_first(ops):
    """The shared-memory opt-in of the large kernels is cached per device ordinal: a process that renders on
    cuda:0 and then on cuda:1 must get the same image from both.  Needs two visible GPUs (skipped on a one-GPU box)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two visible GPUs")
    siren_type, state, z, cam, draws, meta = _full_size_case(1, 32, 8, 16, 80, True)
    imgs = []
    for index in (0, 1):
        d = torch.device("cuda", index)
        from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
        gen = ImplicitGenerator3d(siren_type, 256, 32, 4, 256)
        gen.load_state_dict(state, strict=True)
        gen = gen.to(d)
        gen.set_device(d)
        with torch.no_grad():
            px, _ = gen((z[0].to(d), z[1].to(d)), cam.to(d), draws={k: v.to(d) for k, v in draws.items()}, **meta)
        imgs.append(px.cpu())
    assert torch.equal(imgs[0], imgs[1])


def _dev_tree(z):
    if isinstance(z, (list, tuple)):
        return type(z)(_dev_tree(t) for t in z)
    return dev(z)


@pytest.mark.parametrize("name", LIBRARY_FIXTURES)
def test_library_mlp_decoders_forward_vs_reference(name):
    """TALLSIREN (per-point FiLM, configs/thousand/direct_volume/indirect.py), TALLSIREN_dgx, SHORTSIREN_FG_Pyrmd: the library's
    ray / gather / compositing kernels around a PyTorch MLP (generators/siren_library.py) against fixtures recorded from the
    reference classes; also the secondary boundary and staged_forward."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    siren_type, state, z, cam, draws, meta, taps, (z_dim, input_dim) = library_fixture_inputs(name)
    gen = ImplicitGenerator3d(siren_type, z_dim, input_dim, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda").eval()
    gen.set_device(torch.device("cuda"))
    zc, d = _dev_tree(z), {k: dev(v) for k, v in draws.items()}
    B, img, S = cam.shape[0], meta["img_size"], meta["num_steps"]
    with torch.no_grad():
        pixels, depth = gen(zc, dev(cam), draws=d, **meta)
        rs = gen.siren(dev(taps["points_coarse"].reshape(B, -1, 3)), zc, img, S)
        px2, _ = gen.staged_forward(zc, dev(cam), max_batch_size=1, draws=None, **dict(meta, nerf_noise=0.0))
    err_mlp = (rs.cpu().reshape(B, -1, S, 4) - taps["rgb_sigma_coarse"]).abs().max().item()
    psnr = oracle.psnr(pixels.cpu(), taps["pixels"])
    print(f"{name}: MLP max-abs {err_mlp:.2e}, pixels max-abs {(pixels.cpu() - taps['pixels']).abs().max().item():.2e}, PSNR {psnr:.1f} dB")
    assert err_mlp < 5e-4 and psnr >= 60.0
    assert torch.allclose(depth.cpu(), taps["depth"], rtol=0, atol=2e-3)
    assert px2.shape == pixels.shape and torch.isfinite(px2).all()


def test_fp16_host_volume_path(ops):
    """A feature volume held in fp16 (the encoder's autocast dtype; what bench.py's end-to-end leg uploads): the layout pass widens
    it exactly (cng_volume_f16_to_channels_last), and the image stays at the tensor-core bar against the oracle on the fp32 volume."""
    B, img, S, V = 1, 64, 12, 32
    siren_type, state, z, cam, draws, meta = _full_size_case(B, img, S, V, 60, True)
    v16 = dev(z[0]).half()
    assert torch.equal(ops.volume_to_channels_last(v16), ops.volume_to_channels_last(v16.float()))
    cl3 = v16.contiguous(memory_format=torch.channels_last_3d)          # 16-bit channels-last (U-Net under autocast): no layout kernel
    n0 = ops.launch_count
    assert torch.equal(ops.volume_to_channels_last(cl3), ops.volume_to_channels_last(v16))
    assert ops.launch_count - n0 == 1
    ref = oracle.render(state, siren_type, z, cam, draws, **meta)
    d = {k: dev(v) for k, v in draws.items()}
    for precision, min_psnr in (("fp32", 50.0), ("bf16", 40.0)):
        gen = _generator(siren_type, state, precision)
        with torch.no_grad():
            px, dp = gen((v16, dev(z[1])), dev(cam), draws=d, **meta)
        psnr = oracle.psnr(px.cpu(), ref["pixels"])
        print(f"fp16 host volume, {precision} MLP: PSNR {psnr:.1f} dB vs the oracle on the fp32 volume")
        assert psnr >= min_psnr


def test_config4_frame_vs_oracle():
    """BASELINE configs[3]: one frame of the video workload at full size -- 256x256, 48+48 samples, 64^3 volume -- through
    staged_forward (chunked), against the oracle on the same draws (~15 s of CPU)."""
    B, img, S, V = 1, 256, 48, 64
    siren_type, state, z, cam, draws, meta = _full_size_case(B, img, S, V, 50, True)
    ref = oracle.render(state, siren_type, z, cam, draws, **meta)
    sigma_gain = oracle.DENSE_HEAD_GAINS[siren_type][0]
    d = {k: dev(v) for k, v in draws.items()}
    zc = (dev(z[0]), dev(z[1]))
    gen = _generator(siren_type, state, "bf16")
    with torch.no_grad():
        out = gen._render(zc[0], zc[1], dev(cam), img, FOV, 0.25, 1.95, S, True, dict(meta, draws=d), taps=True)
        px, dp = gen.staged_forward(zc, dev(cam), max_batch_size=1, draws=d, **meta)
    assert torch.equal(px, out["pixels"]) and torch.equal(dp, out["depth"])
    _check_indexing_bit_exact(out, draws, S, 0)
    d_c = (out["rgb_sigma_coarse"].cpu() - ref["rgb_sigma_coarse"]).abs()
    psnr = oracle.psnr(px.cpu(), ref["pixels"])
    print(f"config 4 frame bf16: PSNR {psnr:.1f} dB whole image, coarse rgb err {d_c[..., :3].max().item():.2e}, sigma err {d_c[..., 3].max().item():.2e}, "
          f"depth err mean {(dp.cpu() - ref['depth']).abs().mean().item():.2e}")
    assert d_c[..., :3].max().item() < 1e-2 and d_c[..., 3].max().item() < 1e-2 * sigma_gain
    assert psnr >= 40.0


def test_film_parameters_kernel(ops):
    """a5: mapping network on the library's own kernel: matches the oracle and does not depend on the batch size."""
    st = oracle.init_generator_state("TALLSIREN_FG", seed=3)
    g = torch.Generator().manual_seed(4)
    glob = torch.randn((8, 256), generator=g) * 0.05 + 0.19
    w, b = st["siren.mapping_network.weight"], st["siren.mapping_network.bias"]
    freq, phase = ops.film_parameters(dev(glob), dev(w), dev(b))
    f_ref, p_ref = oracle.film_parameters(glob, w, b)
    assert freq.shape == f_ref.shape and phase.shape == p_ref.shape
    assert torch.allclose(freq.cpu(), f_ref, rtol=0, atol=2e-5) and torch.allclose(phase.cpu(), p_ref, rtol=0, atol=2e-6)
    f1, p1 = ops.film_parameters(dev(glob[3:4]), dev(w), dev(b))
    assert torch.equal(f1[0], freq[3]) and torch.equal(p1[0], phase[3])


def test_sample_generator_sigma_grid():
    """extract_shapes.sample_generator (extract_shapes.py:40-78): dense sigma grid through gen.siren, chunked."""
    from conditioned_nerf_gan_b200 import extract_shapes
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs("fwd_TALLSIREN_FG")
    gen = _generator(siren_type, state, "fp32")
    N = 20
    z1 = (z[0][:1], z[1][:1])
    grid = extract_shapes.sample_generator(gen, dev_z(z1), voxel_resolution=N, cube_length=1.2, max_points=3000)
    assert grid.shape == (N, N, N) and grid.dtype == np.float32
    pts, _, _ = oracle.dense_grid_samples(N, (0, 0, 0), 1.2)
    vol = z1[0]
    feat = torch.nn.functional.grid_sample(vol, (pts / 0.6).reshape(1, 1, 1, -1, 3), mode="bilinear", align_corners=False,
                                           padding_mode="border").reshape(1, 32, -1).permute(0, 2, 1)
    spec, ws, bs = oracle._split_state(state, siren_type)
    freq, phase = oracle.film_parameters(z1[1], state["siren.mapping_network.weight"], state["siren.mapping_network.bias"])
    ref = oracle.film_siren_mlp(feat, ws, bs, freq, phase, state["siren.final_layer.weight"], state["siren.final_layer.bias"], True)
    assert np.abs(grid.reshape(-1) - ref[0, :, 3].numpy()).max() < 5e-4


def test_inference_drivers():
    """generate_img (utils.py:60-82) and the video frame loop (inference.py:441-486) on staged_forward."""
    from conditioned_nerf_gan_b200 import inference
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs("fwd_DOUBLESIREN_FG")
    gen = _generator(siren_type, state, "fp32")
    meta = dict(meta, nerf_noise=0.0, cam_r_start=0.9, cam_r_end=1.3, hierarchical_sample=True)
    img, depth3 = inference.generate_img(gen, dev_z(z), dev(cam), {k: v for k, v in meta.items() if not k.startswith("cam_r")})
    B = cam.shape[0]
    assert img.device.type == "cpu" and img.shape == (B, 3, meta["img_size"], meta["img_size"])
    assert depth3.shape == (B, 3, meta["img_size"], meta["img_size"]) and torch.equal(depth3[:, 0], depth3[:, 2])
    z1 = (dev(z[0][:1]), dev(z[1][:1]))
    torch.manual_seed(0)
    frames = inference.render_video_frames(gen, z1, meta, num_frames=8, fps=2, max_batch_size=3)
    assert frames.shape == (8, 3, meta["img_size"], meta["img_size"]) and frames.device.type == "cpu" and torch.isfinite(frames).all()
    # frame k equals a direct forward with that pose and that fov
    cam2world, fov = inference.video_camera_path(8, 2, 0.9, 1.3, "y", "cuda")
    R, S = meta["img_size"] ** 2, meta["num_steps"]
    d = {"u_jitter": torch.rand((1, R, S, 1), device="cuda"), "noise_coarse": torch.zeros((1, R, S, 1), device="cuda"),
         "u_resample": torch.rand((R, S), device="cuda"), "noise_final": torch.zeros((1, R, 2 * S, 1), device="cuda")}
    m = {kk: v for kk, v in meta.items() if kk not in ("fov", "cam_r_start", "cam_r_end")}
    with torch.no_grad():
        a, _ = gen(z1, cam2world[5:6], fov=float(fov[5]), draws=d, **m)
        b, _ = gen.staged_forward(z1, cam2world[5:6], fov=[float(fov[5])], draws=d, **m)
    assert torch.equal(a, b)


def test_cuda_graph_replay_matches_eager():
    """graphs.GraphedRender: the captured forward replays to the same image as the eager call (replayed draws), also
    after the inputs change."""
    from conditioned_nerf_gan_b200.graphs import GraphedRender
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs("fwd_TALLSIREN_FG")
    gen = _generator(siren_type, state, "bf16")
    d = {k: dev(v) for k, v in draws.items()}
    zc, camc = dev_z(z), dev(cam)
    render = GraphedRender(gen, zc, camc, draws=d, **meta)
    with torch.no_grad():
        a, da = gen(zc, camc, draws=d, **meta)
    b, db = render(zc, camc)
    assert torch.equal(a, b) and torch.equal(da, db)
    z2 = (zc[0] * 0.5, zc[1] + 0.01)
    cam2 = camc.flip(0).contiguous()
    with torch.no_grad():
        a2, da2 = gen(z2, cam2, draws=d, **meta)
    b2, db2 = render(z2, cam2)
    assert torch.equal(a2, b2) and torch.equal(da2, db2) and not torch.equal(a2, a)
    # without replayed draws the captured torch.rand / randn advance with every replay
    render2 = GraphedRender(gen, zc, camc, **dict(meta, nerf_noise=0.0))
    p1 = render2(zc, camc)[0].clone()
    p2 = render2(zc, camc)[0].clone()
    assert torch.isfinite(p1).all() and not torch.equal(p1, p2)
