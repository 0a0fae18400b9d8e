"""Parity of the backward kernels against torch autograd through the CPU oracle.
B200 only (-m gpu).  fp32 stages: <= 1e-4 relative; the MLP backward recomputes with bf16 GEMM operands
(like the forward tensor-core path), so its gradients are compared by relative L2 error and cosine."""
import numpy as np
import pytest
import torch

from conftest import LIBRARY_FIXTURES, fixture_inputs, latent_fixture_inputs, library_fixture_inputs
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


# ------------------------------------------------------------------------------------------------
# a14, MLP part: the tcgen05 dgrad chain and split-K weight gradient on synthetic dumps (formats: include/cng_b200.h)
# ------------------------------------------------------------------------------------------------
def to_tile_images(x, dtype):
    """[P, 64*nb] float -> uint8 [T, nb*16384]: per 128-point tile nb K-blocks of [128 rows][64 x 16 bit], 128-byte swizzle."""
    P, W = x.shape
    nb, T = W // 64, (P + 127) // 128
    xp = torch.zeros((T * 128, W), dtype=torch.float32)
    xp[:P] = x
    x16 = xp.to(dtype).view(torch.int16).view(T, 128, nb, 8, 8)          # [tile, row, block, 16-byte chunk, element]
    out = torch.empty((T, nb, 128, 8, 8), dtype=torch.int16)
    rows = torch.arange(128)
    for c in range(8):
        out[:, :, rows, c ^ (rows & 7), :] = x16[:, :, :, c, :].permute(0, 2, 1, 3)
    return out.view(torch.uint8).reshape(T, nb * 16384)


def from_tile_images(img, P, dtype, nb=4):
    T = img.shape[0]
    v = img.cpu().contiguous().view(torch.int16).view(T, nb, 128, 8, 8)
    out = torch.empty((T, 128, nb, 8, 8), dtype=torch.int16)
    rows = torch.arange(128)
    for c in range(8):
        out[:, :, :, c, :] = v[:, :, rows, c ^ (rows & 7), :].permute(0, 2, 1, 3)
    return out.view(dtype).reshape(T * 128, nb * 64)[:P].float()


def g_codes(g):
    """cos -> the 8-bit code of the dump, round(127 g) + 128 (round half to even, as the kernel's magic-number add does)."""
    return (torch.round(g.float() * 127.0) + 128).to(torch.uint8)


def g_dequant(g, bits):
    """The value the dgrad chain multiplies with: fp16 rounding of cos, or (code - 128) / 127."""
    return g.to(torch.float16).double() if bits == 16 else (g_codes(g).double() - 128) / 127


def to_g_images(g, bits):
    """[P, 256] float -> uint8 [T, 128 * 256 * bits / 8] in the epilogue's register order: fp16 [cc 8][q 4][i 4][lane 32][8], or
    8-bit codes [cc 8][q 4][h 2][lane 32][16]."""
    P = g.shape[0]
    T = (P + 127) // 128
    if bits == 16:
        gp = torch.zeros((T * 128, 256), dtype=torch.float32)
        gp[:P] = g
        v = gp.to(torch.float16).view(T, 4, 32, 8, 4, 8)                 # [tile, q, lane, cc, i, e]
        return v.permute(0, 3, 1, 4, 2, 5).contiguous().view(torch.uint8).reshape(T, 65536)
    gp = torch.full((T * 128, 256), 128, dtype=torch.uint8)
    gp[:P] = g_codes(g)
    v = gp.view(T, 4, 32, 8, 2, 16)                                      # [tile, q, lane, cc, h, e]
    return v.permute(0, 3, 1, 4, 2, 5).contiguous().reshape(T, 32768)


def from_g_images(img, P, bits):
    """uint8 [T, G] -> [P, 256] float."""
    T = img.shape[0]
    if bits == 16:
        v = img.cpu().contiguous().view(torch.float16).view(T, 8, 4, 4, 32, 8)       # [tile, cc, q, i, lane, e]
        return v.permute(0, 2, 4, 1, 3, 5).reshape(T * 128, 256)[:P].float()
    v = img.cpu().contiguous().view(T, 8, 4, 2, 32, 16)                  # [tile, cc, q, h, lane, e]
    return (v.permute(0, 2, 4, 1, 3, 5).reshape(T * 128, 256)[:P].float() - 128) / 127


@pytest.fixture(params=[16, 8])
def g_bits(request):
    """Runs a test in both formats of the cos(u) dump (cng_internal_set_g_dump_bits; 16 is the default)."""
    import ctypes
    from conditioned_nerf_gan_b200 import _lib
    lib = _lib.load()
    lib.cng_internal_set_g_dump_bits.argtypes = [ctypes.c_int]
    lib.cng_internal_set_g_dump_bits.restype = None
    lib.cng_internal_set_g_dump_bits(request.param)
    yield request.param
    lib.cng_internal_set_g_dump_bits(0)


@pytest.mark.parametrize("x_dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("P,L", [(1000, 3), (128, 1), (128 * 301 + 5, 2), (77, 4)])
def test_wgrad_kernel_vs_torch(ops, x_dtype, P, L):
    """cng_film_siren_wgrad: dW_l = dz_l^T x_l by tcgen05 with MN-major operands read from the tile images (dz bf16; x bf16, or
    fp16 converted to bf16 in shared memory), fp32 TMEM accumulation, red.global.add flush; column sums of dz."""
    g = torch.Generator().manual_seed(P + L)
    dz = [torch.randn((P, 256), generator=g) * (0.5 + l) for l in range(L)]
    xs = [torch.sin(torch.randn((P, 256), generator=g) * 3) for _ in range(max(L - 1, 1))]
    feat = torch.randn((P, 32), generator=g) * 0.3
    hi = feat.to(x_dtype).float()
    lo = (feat - hi).to(x_dtype).float()
    dz_img = torch.stack([to_tile_images(d, torch.bfloat16) for d in dz]).cuda()
    x_img = torch.stack([to_tile_images(x, x_dtype) for x in xs]).cuda()
    f_img = to_tile_images(torch.cat([hi, lo], dim=1), x_dtype).cuda()
    dW = [torch.zeros((256, 32 if l == 0 else 256), device="cuda") for l in range(L)]
    colsum = torch.zeros((L, 256), device="cuda")
    for rep in range(2):                                                  # accumulation semantics: the second call doubles everything
        ops.film_siren_wgrad(dz_img, x_img, f_img, P, L, x_dtype == torch.float16, dW, colsum)
    torch.cuda.synchronize()
    for l in range(L):
        dzr = dz[l].to(torch.bfloat16).double()
        # an fp16 dump is rounded to bf16 element by element before the MMA
        rb = (lambda t: t.to(torch.bfloat16).double()) if x_dtype == torch.float16 else (lambda t: t.double())
        xr = (rb(hi) + rb(lo)) if l == 0 else rb(xs[l - 1].to(x_dtype).float())
        ref = 2 * dzr.t() @ xr
        e = rel_l2(dW[l].cpu(), ref)
        print(f"wgrad P={P} L={L} {x_dtype} layer {l}: rel-L2 {e:.2e}")
        assert e < 1e-5, (l, e)
        assert rel_l2(colsum[l].cpu(), 2 * dzr.sum(0)) < 1e-5


@pytest.mark.parametrize("P,L,sig", [(1000, 3, True), (128, 1, False), (128 * 301 + 5, 2, True), (77, 8, False)])
def test_dgrad_kernel_vs_torch(ops, P, L, sig, g_bits):
    """cng_film_siren_dgrad: the fused chain d_o -> dy -> dz_l = dy * g_l -> dy = dz_l W_l ... -> d_feat on synthetic g, in both dump formats."""
    g = torch.Generator().manual_seed(P * 3 + L)
    ws = [torch.randn((256, 32 if l == 0 else 256), generator=g) * (0.2 if l == 0 else 0.06) for l in range(L)]
    fw = torch.randn((4, 256), generator=g) * 0.1
    gs = [torch.cos(torch.randn((P, 256), generator=g) * 3) for _ in range(L)]
    gs[0][0, :4] = torch.tensor([1.0, -1.0, 0.0, 0.5 / 127])                    # the ends of the code range, zero, a rounding tie
    d_out = torch.randn((P, 4), generator=g)
    out = torch.rand((P, 4), generator=g)
    wt = ops.film_siren_wt_images([w.cuda() for w in ws], fw.cuda())
    assert ops.g_image_bytes() == 128 * 256 * g_bits // 8
    g_img = torch.stack([to_g_images(x, g_bits) for x in gs]).cuda()
    d_fb = torch.zeros((4,), device="cuda")
    d_feat, dz_img = ops.film_siren_dgrad(d_out.cuda(), out.cuda(), sig, L, wt, g_img, d_fb)
    torch.cuda.synchronize()
    bf = lambda t: t.to(torch.bfloat16).double()
    d_o = d_out.clone().double()
    if sig:
        d_o[:, :3] *= (out[:, :3] * (1 - out[:, :3])).double()
    assert rel_l2(d_fb.cpu(), d_o.sum(0)) < 1e-5
    dy = d_o @ bf(fw)                       # d_o enters as hi + lo (~fp32), Wf as bf16
    for l in reversed(range(L)):
        dz = dy * g_dequant(gs[l], g_bits)
        got = from_tile_images(dz_img[l], P, torch.bfloat16)
        e = rel_l2(got, dz)
        print(f"dgrad P={P} L={L} layer {l}: dz rel-L2 {e:.2e}")
        assert e < 6e-3, (l, e)             # bf16 rounding of dz (2^-9 relative per element), accumulated over the layers
        dy = bf(dz.float()) @ bf(ws[l])
    e = rel_l2(d_feat.cpu(), dy)
    print(f"dgrad P={P} L={L}: d_feat rel-L2 {e:.2e}")
    assert e < 6e-3 and d_feat.shape == (P, 32)


def test_fwd_train_dumps_vs_oracle(ops, g_bits):
    """cng_film_siren_fwd_train: x tile images, g = cos(u) and the layer-0 operand block against the oracle's activations."""
    from test_gpu_parity import _mlp_setup
    B, N = 2, 300
    spec, ws, bs, feat, freq, phase, fw, fb, ref = _mlp_setup("SHORTSIREN_FG", B, N, 0.3)
    out, xs, gs, fd = ops.film_siren_fwd_train(dev(feat), [dev(w) for w in ws], [dev(b) for b in bs], dev(freq), dev(phase), dev(fw), dev(fb),
                                               spec["sigmoid_rgb"], "fp16")
    torch.cuda.synchronize()
    assert (out.cpu() - ref).abs().max().item() < 1e-2
    L, tpi = len(ws), (N + 127) // 128
    x = feat
    for l in range(L):
        u = freq[:, l * 256:(l + 1) * 256].unsqueeze(1) * torch.nn.functional.linear(x, ws[l], bs[l]) + phase[:, l * 256:(l + 1) * 256].unsqueeze(1)
        x = torch.sin(u)
        gref = torch.cos(u)
        for b in range(B):
            got_x = from_tile_images(xs[l, b * tpi:(b + 1) * tpi], N, torch.float16)
            assert (got_x - x[b]).abs().max().item() < 5e-2, (l, b)        # hidden activations of SHORTSIREN_FG carry the fp16-operand error of the layers before
            got_g = from_g_images(gs[l, b * tpi:(b + 1) * tpi], N, g_bits)
            assert (got_g - gref[b]).abs().max().item() < 0.3, (l, b, (got_g - gref[b]).abs().max().item())     # u carries the 16-bit operand error x freq ~ 30
    for b in range(B):
        f = from_tile_images(fd[b * tpi:(b + 1) * tpi], N, torch.float16, nb=1)
        assert (f[:, :32] + f[:, 32:] - feat[b]).abs().max().item() < 1e-5


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


def test_latent_shortsiren_backward_vs_oracle_autograd():
    """Gradients of the position-input SHORTSIREN w.r.t. every parameter (mapping network included) and the latent vector."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    state, latent, cam, draws, meta, _ = latent_fixture_inputs()
    B, img = cam.shape[0], meta["img_size"]
    g = torch.Generator().manual_seed(2)
    d_pix, d_dep = torch.randn((B, 3, img, img), generator=g), torch.randn((B, img, img), generator=g)
    st = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    z_r = latent.clone().requires_grad_(True)
    out = oracle.render_with_grad(st, "SHORTSIREN", z_r, cam, draws, **meta)
    ((out["pixels"] * d_pix).sum() + (out["depth"] * d_dep).sum()).backward()
    gen = ImplicitGenerator3d("SHORTSIREN", latent.shape[1], 3, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    gen.siren.precision = "fp32"
    z_d = dev(latent).requires_grad_(True)
    pixels, depth = gen(z_d, dev(cam), draws={k: dev(v) for k, v in draws.items()}, **meta)
    ((pixels * dev(d_pix)).sum() + (depth * dev(d_dep)).sum()).backward()
    pairs = [("latent", z_d.grad, z_r.grad)] + [(k, p.grad, st["siren." + k].grad) for k, p in gen.siren.named_parameters()]
    for k, got, ref in pairs:
        assert got is not None and got.shape == ref.shape, k
        e, c = rel_l2(got.cpu(), ref), cosine(got.cpu(), ref)
        print(f"  SHORTSIREN {k}: rel-L2 {e:.3e} cos {c:.6f}")
        assert c > 0.9995 and e < 2e-2, (k, e, c)


@pytest.mark.parametrize("name", ["fwd_TALLSIREN_FG", "fwd_TALLSIREN_dRes"])
def test_kept_dumps_backward_equals_recompute_backward(name):
    """Training with the forward's dumps kept for the backward (CNG_KEEP_DUMPS: one training-mode forward for the whole batch -- or
    for its first k items when memory holds only a part of it --, the dgrad / weight-gradient kernels reading each item's tiles out
    of the batch dump) against the recomputing backward."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d, autograd
    state, siren_type, z, cam, draws, meta, _ = fixture_inputs(name)
    B, img = cam.shape[0], meta["img_size"]
    g = torch.Generator().manual_seed(3)
    d_pix, d_dep = dev(torch.randn((B, 3, img, img), generator=g)), dev(torch.randn((B, img, img), generator=g))
    gen = ImplicitGenerator3d(siren_type, 32 if "dRes" in siren_type else 256, 32, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    gen.siren.precision = "fp16"
    film = isinstance(z, tuple)
    grads = {}
    prev = autograd.KEEP_DUMPS
    try:
        for mode in ("0", "1", "first:1"):
            autograd.KEEP_DUMPS = mode
            gen.zero_grad(set_to_none=True)
            vol = dev(z[0] if film else z).requires_grad_(True)
            glob = dev(z[1]).requires_grad_(True) if film else None
            pixels, depth = gen((vol, glob) if film else vol, dev(cam), draws={k: dev(v) for k, v in draws.items()}, **meta)
            ((pixels * d_pix).sum() + (depth * d_dep).sum()).backward()
            grads[mode] = {"volume": vol.grad.clone(), **{k: p.grad.clone() for k, p in gen.siren.named_parameters()}}
            if film:
                grads[mode]["global"] = glob.grad.clone()
    finally:
        autograd.KEEP_DUMPS = prev
    assert B >= 2, "the partial mode needs a fixture with at least two items"
    worst = 0.0
    for k in grads["0"]:
        e = max(rel_l2(grads["1"][k].cpu(), grads["0"][k].cpu()), rel_l2(grads["first:1"][k].cpu(), grads["0"][k].cpu()))
        worst = max(worst, e)
        print(f"  kept / partly kept vs recompute {name} {k}: rel-L2 {e:.2e}")
    # same backward kernels on the same dumps; the two forwards differ in which sines take the FMA-pipe polynomial (1 in 8 vs 1 in 4,
    # 7e-5 each), which the FiLM frequencies (~30 per layer) amplify into the loss gradient
    assert worst < 1e-2, worst


@pytest.mark.parametrize("name", LIBRARY_FIXTURES)
def test_library_mlp_decoders_backward_vs_oracle_autograd(name):
    """Gradients of the library-MLP decoders (checkpointed PyTorch MLP between the library's gather / scatter and compositing
    kernels) w.r.t. every parameter, every feature volume and the global feature."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    siren_type, state, z, cam, draws, meta, _, (z_dim, input_dim) = library_fixture_inputs(name)
    B, img = cam.shape[0], meta["img_size"]
    g = torch.Generator().manual_seed(4)
    d_pix, d_dep = torch.randn((B, 3, img, img), generator=g), torch.randn((B, img, img), generator=g)

    def leaves(zz, to_dev):
        if isinstance(zz, (list, tuple)):
            return type(zz)(leaves(t, to_dev) for t in zz)
        return (dev(zz) if to_dev else zz.clone()).requires_grad_(True)

    def flat(zz):
        return [t for x in zz for t in flat(x)] if isinstance(zz, (list, tuple)) else [zz]

    st = {k: v.clone().requires_grad_(True) for k, v in state.items()}
    z_r = leaves(z, False)
    out = oracle.render_with_grad(st, siren_type, z_r, cam, draws, **meta)
    ((out["pixels"] * d_pix).sum() + (out["depth"] * d_dep).sum()).backward()
    gen = ImplicitGenerator3d(siren_type, z_dim, input_dim, 4, 256)
    gen.load_state_dict(state, strict=True)
    gen = gen.to("cuda")
    z_d = leaves(z, True)
    pixels, depth = gen(z_d, dev(cam), draws={k: dev(v) for k, v in draws.items()}, **meta)
    ((pixels * dev(d_pix)).sum() + (depth * dev(d_dep)).sum()).backward()
    pairs = [(f"z[{i}]", a.grad, b.grad) for i, (a, b) in enumerate(zip(flat(z_d), flat(z_r)))]
    pairs += [(k, p.grad, st["siren." + k].grad) for k, p in gen.siren.named_parameters()]
    for k, got, ref in pairs:
        assert got is not None and got.shape == ref.shape, k
        e, c = rel_l2(got.cpu(), ref), cosine(got.cpu(), ref)
        assert c > 0.9995 and e < 2e-2, (k, e, c)
    print(f"{name}: worst rel-L2 {max(rel_l2(a.cpu(), b) for _, a, b in pairs):.2e}")


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
