"""The callers either side of the rendering path (BASELINE config 3): 3D U-Net encoder,
progressive discriminator and the GAN train step, against tests/golden/train_step.npz -- recorded by
tests/golden/make_golden.py from the REFERENCE's own unet3d.UNet3D, ProgressiveDiscriminator and ImplicitGenerator3d."""
import pytest
import torch

from conftest import load_golden
from oracle import nerf_path as oracle
from oracle import train_step as ts


def _modules():
    from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator
    from conditioned_nerf_gan_b200.generators.unet3d import UNet3D
    enc, disc = UNet3D(**ts.TINY_UNET), ProgressiveDiscriminator()
    ts.fill_params(enc, 1)
    ts.fill_params(disc, 2)
    return enc, disc


def test_state_dict_keys_are_the_reference_layout():
    enc, disc = _modules()
    ek, dk = list(enc.state_dict()), list(disc.state_dict())
    assert ek[:3] == ["encoders.0.basic_module.SingleConv1.groupnorm.weight", "encoders.0.basic_module.SingleConv1.groupnorm.bias",
                      "encoders.0.basic_module.SingleConv1.conv.weight"]
    assert ek[-2:] == ["final_conv.weight", "final_conv.bias"] and any(k.startswith("decoders.0.basic_module.SingleConv2.") for k in ek)
    assert dk[:4] == ["layers.0.network.0.conv.weight", "layers.0.network.0.conv.bias", "layers.0.network.2.conv.weight", "layers.0.network.2.conv.bias"]
    assert "layers.0.proj.weight" in dk and "fromRGB.8.model.0.bias" in dk and dk[-2:] == ["final_layer.weight", "final_layer.bias"]
    assert disc.state_dict()["layers.0.network.0.conv.weight"].shape == (32, 18, 3, 3)     # 16 + 2 coordinate channels
    # the full-size encoder of configs/thousand/special.py:53-62
    from conditioned_nerf_gan_b200.generators.unet3d import UNet3D
    full = UNet3D(in_channels=4, out_channels=32, f_maps=32, num_levels=4, is_segmentation=False, final_sigmoid=False, return_global=True)
    assert 4.0e6 < sum(p.numel() for p in full.parameters()) < 4.2e6          # 4.08 M parameters


def test_unet_and_discriminator_match_reference_outputs():
    fx, _ = load_golden("train_step")
    enc, disc = _modules()
    sample = ts.tiny_sample()
    with torch.no_grad():
        fv, glob = enc(sample["voxel"])
        assert fv.shape == fx["unet/volume"].shape and fv.is_contiguous(memory_format=torch.channels_last_3d)
        torch.testing.assert_close(fv.contiguous(), fx["unet/volume"], rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(glob, fx["unet/global"], rtol=1e-4, atol=1e-6)
        for size in (16, 64):
            torch.testing.assert_close(disc(fx[f"disc/in{size}"], 0.3), fx[f"disc/out{size}"], rtol=1e-4, atol=1e-6)


def test_pyramid_and_residual_unets_match_reference_outputs():
    """PyramidUNet3D (every decoder level, no final convolution) and ResidualUNet3D (ExtResNetBlock, transposed-convolution
    upsampling, summation joining), generators/unet3d.py:829-898, against outputs of the reference's classes."""
    from conditioned_nerf_gan_b200.generators.unet3d import PyramidUNet3D, ResidualUNet3D
    fx, _ = load_golden("train_step")
    voxel = ts.tiny_sample()["voxel"]
    pyr = PyramidUNet3D(**dict(ts.TINY_UNET, num_levels=3))
    res = ResidualUNet3D(**dict(ts.TINY_UNET, num_levels=3, return_global=False, out_channels=16))
    ts.fill_params(pyr, 3)
    ts.fill_params(res, 4)
    assert not any(k.startswith("final_conv") for k in pyr.state_dict())
    assert "decoders.0.upsampling.upsample.weight" in res.state_dict() and "encoders.1.basic_module.conv3.groupnorm.weight" in res.state_dict()
    with torch.no_grad():
        levels, glob = pyr(voxel)
        assert len(levels) == 2
        for i, lv in enumerate(levels):
            torch.testing.assert_close(lv.contiguous(), fx[f"unet_pyramid/level{i}"], rtol=1e-4, atol=2e-5)
        torch.testing.assert_close(glob, fx["unet_pyramid/global"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(res(voxel).contiguous(), fx["unet_residual/volume"], rtol=1e-4, atol=2e-5)


def test_curriculum_helpers_match_reference():
    """extract_metadata / last_upsample_step / next_upsample_step (configs/curriculums.py:83-137) on a four-stage curriculum,
    against values computed by the reference's own functions; GanTrainStep.set_alpha reads the stage start from them."""
    from conditioned_nerf_gan_b200 import curriculums
    fx, _ = load_golden("train_step")
    cur = ts.tiny_curriculum()
    for step, (img, batch, last, nxt) in fx["curriculum/json"].items():
        step = int(step)
        md = curriculums.extract_metadata(cur, step)
        assert (md["img_size"], md["batch_size"], md["fade_steps"]) == (img, batch, 2000)
        assert curriculums.last_upsample_step(cur, step) == last
        assert min(curriculums.next_upsample_step(cur, step), 1e9) == nxt
    assert curriculums.update_recursive({"a": {"b": 1}}, {"a": {"c": 2}, "d": 3}) == {"a": {"b": 1, "c": 2}, "d": 3}


def test_discriminator_variants_match_reference_outputs():
    """ProgressiveEncoderDiscriminator (prediction + latent + position heads) and ProgressiveDiscriminator_inputCat (condition
    image concatenated to the input), discriminators.py:202-335."""
    from conditioned_nerf_gan_b200.discriminators import ProgressiveDiscriminator_inputCat, ProgressiveEncoderDiscriminator
    fx, _ = load_golden("train_step")
    img = fx["disc/in16"]
    enc, cat = ProgressiveEncoderDiscriminator(), ProgressiveDiscriminator_inputCat()
    ts.fill_params(enc, 5)
    ts.fill_params(cat, 5)
    with torch.no_grad():
        pred, latent, position = enc(img, 0.3)
        assert pred.shape == (2, 1) and latent.shape == (2, 256) and position.shape == (2, 2)
        torch.testing.assert_close(torch.cat([pred, latent, position], dim=1), fx["disc_enc/out16"], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(cat(img, 0.3, cond=img.flip(0)), fx["disc_cat/out16"], rtol=1e-4, atol=1e-6)


def test_discriminator_ignores_curriculum_keys_and_fades_in():
    _, disc = _modules()
    img = torch.rand((1, 3, 32, 32)) * 2 - 1
    a = disc(img, 1.0, cond=None, img_size=32, fov=30, batch_split=2)
    b = disc(img, 0.0)
    assert a.shape == (1, 1) and not torch.allclose(a, b)
    with pytest.raises(KeyError):
        disc(torch.rand((1, 3, 24, 24)), 1.0)          # img_size_to_layer lookup, discriminators.py:185-187


def test_train_step_harness_with_our_encoder_and_discriminator_matches_reference_modules():
    """oracle.RefTrainStep (utils.py:621-842 restated) around THIS repository's U-Net / discriminator and the oracle's
    renderer reproduces the losses and gradient norms recorded with the reference's three modules."""
    fx, _ = load_golden("train_step")
    enc, disc = _modules()
    gen = ts.OracleGenerator(ts.TINY_SIREN, oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0))
    harness = ts.RefTrainStep(gen, enc, disc, dict(ts.tiny_config(), draws=ts.tiny_draws()), alpha=0.3)
    sample = ts.tiny_sample()
    for i in range(2):
        rec = harness.step(sample)
        for k, v in rec.items():
            want = float(fx[f"step{i}/{k}"])
            assert abs(v - want) <= 2e-3 * abs(want) + 1e-5, (i, k, v, want)


def test_step_shares_one_encoder_forward_without_changing_the_update():
    """GanTrainStep.step() encodes the voxels once for the D step and the G/E step (the reference does it twice with unchanged
    encoder weights, utils.py:771-775 and :652-655): same losses, gradient norms and updated parameters as the two-pass form."""
    from conditioned_nerf_gan_b200.training import GanTrainStep
    results = []
    for share in (True, False):
        torch.manual_seed(11)
        enc, disc = _modules()
        gen = ts.OracleGenerator(ts.TINY_SIREN, oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0))
        calls = {"n": 0}
        enc.register_forward_hook(lambda *a: calls.__setitem__("n", calls["n"] + 1))
        tr = GanTrainStep(gen, enc, disc, dict(ts.tiny_config(), draws=ts.tiny_draws()), "cpu", amp=False)
        tr.share_encoder_forward = share
        tr.alpha = 0.3
        torch.manual_seed(5)
        for _ in range(2):
            losses = tr.step(ts.tiny_sample())
        assert calls["n"] == (2 if share else 4)
        results.append(({k: float(v) for k, v in losses.items()}, {k: float(v) for k, v in tr.grad_norms.items()},
                        [p.detach().clone() for p in enc.parameters()], [p.detach().clone() for p in gen.parameters()]))
    (l0, n0, e0, g0), (l1, n1, e1, g1) = results
    assert l0 == pytest.approx(l1, rel=1e-6) and n0 == pytest.approx(n1, rel=1e-6)
    for a, b in zip(e0 + g0, e1 + g1):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-8)


def test_gan_train_step_host_logic_matches_reference_on_cpu():
    """training.GanTrainStep itself (optimizers, GradScaler order, clipping, loss bookkeeping) on CPU, with the oracle renderer
    standing in for the CUDA generator: two optimisation steps against the numbers recorded with the reference's modules.
    Also: batch_split = 2 accumulates the two half-batch gradients exactly as the reference's loop does (sum of the two means)."""
    from conditioned_nerf_gan_b200.training import GanTrainStep
    fx, _ = load_golden("train_step")
    enc, disc = _modules()
    gen = ts.OracleGenerator(ts.TINY_SIREN, oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0))
    trainer = GanTrainStep(gen, enc, disc, dict(ts.tiny_config(), draws=ts.tiny_draws()), "cpu", amp=False)
    trainer.alpha = 0.3
    sample = ts.tiny_sample()
    for i in range(2):
        trainer.train_discriminator(sample)
        trainer.train_generator(sample)
        got = {"d_loss": trainer.losses["d_loss"], "g_loss": trainer.losses["g_loss"], "photo_loss": trainer.losses["photo_loss"],
               "norm_D": trainer.grad_norms["D"], "norm_G": trainer.grad_norms["G"], "norm_E": trainer.grad_norms["E"]}
        for k, v in got.items():
            want = float(fx[f"step{i}/{k}"])
            assert abs(float(v) - want) <= 2e-3 * abs(want) + 1e-5, (i, k, float(v), want)
    # batch_split: per-image draws make the two halves independent, so split 2 = sum of the half-batch gradients
    norms = {}
    for splits in (1, 2):
        enc2, disc2 = _modules()
        gen2 = ts.OracleGenerator(ts.TINY_SIREN, oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0))

        class _PerSplitDraws(torch.nn.Module):
            """hands the generator the draws of the images it is asked to render"""
            def __init__(self, inner):
                super().__init__()
                self.inner, self.step, self.calls = inner, 0, 0
            def forward(self, z, cam, **md):
                b, R = cam.shape[0], md["img_size"] ** 2
                lo = (self.calls * b) % ts.TINY_BATCH if b < ts.TINY_BATCH else 0
                self.calls += 1
                d = {k: (v[lo * R:(lo + b) * R] if k == "u_resample" else v[lo:lo + b]) for k, v in ts.tiny_draws().items()}
                return self.inner(z, cam, **dict(md, draws=d))

        tr = GanTrainStep(_PerSplitDraws(gen2), enc2, disc2, dict(ts.tiny_config(), batch_split=splits, enable_discriminator=False), "cpu", amp=False)
        tr.train_generator(sample)
        norms[splits] = (float(tr.grad_norms["G"]), float(tr.grad_norms["E"]), float(tr.losses["photo_loss"]))
    # each split's loss is a mean over half the images: the accumulated gradient is twice the whole-batch-mean gradient
    assert norms[2][0] == pytest.approx(2 * norms[1][0], rel=2e-3) and norms[2][1] == pytest.approx(2 * norms[1][1], rel=2e-3)
    assert norms[2][2] == pytest.approx(norms[1][2], rel=1e-4)



@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.float16, torch.bfloat16])
@pytest.mark.parametrize("N,C,G,shape", [(2, 32, 8, (16, 16, 16)), (3, 4, 1, (9, 10, 11)), (1, 384, 8, (8, 8, 8)), (2, 96, 8, (5, 6, 7)), (2, 16, 8, (12, 12)),
                                         (1, 256, 8, (32, 32, 32)), (2, 6, 3, (7, 5, 3))])
def test_group_norm_channels_last_kernels_vs_torch(dtype, N, C, G, shape):
    """cng_group_norm_fwd / _bwd (the U-Net's GroupNorm on channels-last tensors) against torch's group norm in float64."""
    from conditioned_nerf_gan_b200.generators.unet3d import GroupNormCL
    g = torch.Generator().manual_seed(C * 7 + N)
    x = (torch.randn((N, C, *shape), generator=g) * 1.7 + 0.4)
    fmt = torch.channels_last_3d if len(shape) == 3 else torch.channels_last
    gn = GroupNormCL(G, C)
    with torch.no_grad():
        gn.weight.copy_(torch.randn(C, generator=g) * 0.5 + 1)
        gn.bias.copy_(torch.randn(C, generator=g) * 0.3)
    dy = torch.randn(x.shape, generator=g)
    xr = x.to(dtype).double().requires_grad_(True)
    ref = torch.nn.functional.group_norm(xr, G, gn.weight.detach().double(), gn.bias.detach().double(), gn.eps)
    ref.backward(dy.to(dtype).double())
    gn = gn.cuda()
    xd = x.to(dtype).cuda().contiguous(memory_format=fmt).requires_grad_(True)
    y = gn(xd)
    assert y.dtype == dtype and y.is_contiguous(memory_format=fmt)
    y.backward(dy.to(dtype).cuda().contiguous(memory_format=fmt))
    tol = 2e-5 if dtype == torch.float32 else (2e-3 if dtype == torch.float16 else 1.6e-2)
    assert (y.detach().cpu().double() - ref.detach()).abs().max().item() < tol * 4
    rel = lambda a, b: float((a.double().cpu() - b).norm() / (b.norm() + 1e-30))
    assert rel(xd.grad, xr.grad) < tol, rel(xd.grad, xr.grad)
    # d_weight / d_bias against float64 on the same (rounded) input
    w64, b64 = gn.weight.detach().double().cpu().requires_grad_(True), gn.bias.detach().double().cpu().requires_grad_(True)
    torch.nn.functional.group_norm(x.to(dtype).double(), G, w64, b64, gn.eps).backward(dy.to(dtype).double())
    ew, eb = rel(gn.weight.grad, w64.grad), rel(gn.bias.grad, b64.grad)
    assert ew < tol and eb < tol, (ew, eb)

@pytest.mark.gpu
def test_gan_train_step_matches_reference_on_gpu():
    """The product's GanTrainStep (CUDA rendering path in exact-fp32 mode, cuDNN U-Net / discriminator) against the
    reference-recorded losses and gradient norms of two consecutive optimisation steps."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    from conditioned_nerf_gan_b200.training import GanTrainStep
    fx, _ = load_golden("train_step")
    dev = torch.device("cuda")
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        enc, disc = _modules()
        gen = ImplicitGenerator3d(ts.TINY_SIREN, ts.TINY_ZDIM, 32, 4, 256)
        gen.load_state_dict(oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0), strict=True)
        gen.siren.precision = "fp32"
        enc, disc, gen = enc.to(dev), disc.to(dev), gen.to(dev)
        md = dict(ts.tiny_config(), draws={k: v.to(dev) for k, v in ts.tiny_draws().items()})
        trainer = GanTrainStep(gen, enc, disc, md, dev, amp=False)
        trainer.alpha = 0.3
        sample = {k: v.to(dev) for k, v in ts.tiny_sample().items()}
        for i in range(2):
            trainer.train_discriminator(sample)
            trainer.train_generator(sample)
            got = {"d_loss": trainer.losses["d_loss"], "g_loss": trainer.losses["g_loss"], "photo_loss": trainer.losses["photo_loss"],
                   "norm_D": trainer.grad_norms["D"], "norm_G": trainer.grad_norms["G"], "norm_E": trainer.grad_norms["E"]}
            for k, v in got.items():
                want = float(fx[f"step{i}/{k}"])
                tol = 2e-2 if k.startswith("norm") else 2e-3          # gradient norms: fp32 backward of the MLP recompute (bf16 GEMMs)
                assert abs(float(v) - want) <= tol * abs(want) + 1e-5, (i, k, float(v), want)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev


@pytest.mark.gpu
def test_gan_train_step_amp_full_api():
    """step() under autocast + GradScaler with the tensor-core MLP, random discriminator cameras, batch_split 2: finite
    losses, parameters move, counters advance, the encoder's volume arrives in the gather kernel's layout."""
    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d
    from conditioned_nerf_gan_b200.training import GanTrainStep
    dev = torch.device("cuda")
    enc, disc = _modules()
    gen = ImplicitGenerator3d(ts.TINY_SIREN, ts.TINY_ZDIM, 32, 4, 256)
    gen.load_state_dict(oracle.init_generator_state(ts.TINY_SIREN, z_dim=ts.TINY_ZDIM, seed=0), strict=True)
    enc, disc, gen = enc.to(dev), disc.to(dev), gen.to(dev)
    md = dict(ts.tiny_config(), random_gen_img=True, batch_split=2)
    trainer = GanTrainStep(gen, enc, disc, md, dev, amp=True)
    sample = {k: v.to(dev) for k, v in ts.tiny_sample().items()}
    before = [p.detach().clone() for m in (gen, enc, disc) for p in list(m.parameters())[:2]]
    for _ in range(2):
        losses = trainer.step(sample)
    assert all(torch.isfinite(v).all() for v in losses.values()), losses
    after = [p.detach() for m in (gen, enc, disc) for p in list(m.parameters())[:2]]
    assert any(not torch.equal(a, b) for a, b in zip(before, after))
    assert gen.step == 2 and disc.step == 2 and trainer.metadata["nerf_noise"] == pytest.approx(1.0 - 1 / 5000.0)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16):
        fv, _ = enc(sample["voxel"])
    assert fv.is_contiguous(memory_format=torch.channels_last_3d)


def test_conv2d_r1_double_backward_equals_stock_autograd():
    """conv2d_r1 spells out the convolution's first and second derivative (fast kernels for the R1 penalty): values, gradients
    and the gradient of an R1-style penalty must equal torch's own autograd through F.conv2d (float64, strides 1 and 2)."""
    from conditioned_nerf_gan_b200.discriminators.discriminators import conv2d_r1
    g = torch.Generator().manual_seed(3)
    for stride, padding, k in ((1, 1, 3), (2, 1, 3), (1, 0, 1), (1, 0, 2)):
        x0 = torch.randn((2, 5, 9, 8), generator=g, dtype=torch.float64)
        w0 = torch.randn((7, 5, k, k), generator=g, dtype=torch.float64) * 0.3
        b0 = torch.randn((7,), generator=g, dtype=torch.float64)
        res = []
        for fn in (conv2d_r1, torch.nn.functional.conv2d):
            x, w, b = x0.clone().requires_grad_(True), w0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
            y = torch.tanh(fn(x, w, b, stride, padding))
            gx = torch.autograd.grad(y.sum(), x, create_graph=True)[0]
            loss = y.square().mean() + 0.5 * (gx.reshape(2, -1).norm(2, dim=1) ** 2).mean()
            loss.backward()
            res.append((y.detach(), gx.detach(), x.grad, w.grad, b.grad))
        for a, b_ in zip(*res):
            torch.testing.assert_close(a, b_, rtol=1e-10, atol=1e-12)
