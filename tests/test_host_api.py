"""CPU-only checks of the host-side mirror of the reference's generator interface."""
import pytest
import torch

from conditioned_nerf_gan_b200 import ops
from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d, siren, volumetric_rendering as vr
from oracle import nerf_path as oracle

META = dict(img_size=8, fov=49.134342641202636, ray_start=0.25, ray_end=1.95, num_steps=6,
            hierarchical_sample=True, clamp_mode="relu", nerf_noise=0.0, white_back=True,
            # keys of the curriculum dict the generator must ignore
            batch_size=4, topk_v=0.6, fade_steps=10000, unique_lr=False, betas=(0, 0.9), z_lambda=0)


def make_z(B=1, V=8):
    return torch.zeros((B, 32, V, V, V)), torch.zeros((B, 256))


@pytest.mark.parametrize("name,layers", [("TALLSIREN_FG", 8), ("SHORTSIREN_FG", 4), ("DOUBLESIREN_FG", 2), ("SingleSIREN_dg", 1),
                                         ("TALLSIREN_dg", 8), ("SHORTSIREN_dg", 4), ("DoubleSIREN_dg", 2)])
def test_state_dict_layout_matches_reference(name, layers):
    gen = ImplicitGenerator3d(siren_type=name, z_dim=256, input_dim=32, output_dim=4, hidden_dim=256)
    ref_state = oracle.init_generator_state(name)
    assert set(gen.state_dict().keys()) == set(ref_state.keys())
    for k, v in gen.state_dict().items():
        assert v.shape == ref_state[k].shape, k
    gen.load_state_dict(ref_state, strict=True)
    assert len(gen.siren.network) == layers
    assert gen.epoch == 0 and gen.step == 0
    gen.set_device("cpu")
    assert gen.device == "cpu" and gen.siren.device == "cpu"
    assert gen.siren.training
    gen.eval()
    assert not gen.siren.training


def test_init_distributions():
    torch.manual_seed(0)
    for name, div in (("TALLSIREN_FG", 25.0), ("SHORTSIREN_FG", 12.0)):
        s = ImplicitGenerator3d(name, 256, 32, 4, 256).siren
        assert s.network[0].layer.weight.abs().max() <= 1 / 32            # first_layer_film_sine_init
        bound = (6 / 256) ** 0.5 / div
        assert bound * 0.98 < s.network[1].layer.weight.abs().max() <= bound
        assert s.final_layer.weight.abs().max() <= bound
        assert s.mapping_network.weight.shape == (2 * len(s.network) * 256, 256)


def test_unknown_siren_type_is_attribute_error():
    with pytest.raises(AttributeError):
        ImplicitGenerator3d("NOPE", 256, 32, 4, 256)


def test_missing_curriculum_keys_raise_keyerror():
    gen = ImplicitGenerator3d("DOUBLESIREN_FG", 256, 32, 4, 256)
    cam = torch.eye(4).unsqueeze(0)
    for missing in ("clamp_mode", "nerf_noise"):
        meta = {k: v for k, v in META.items() if k != missing}
        with torch.no_grad(), pytest.raises(KeyError):
            gen(make_z(), cam, **meta)


def test_unknown_clamp_mode_is_type_error():
    gen = ImplicitGenerator3d("DOUBLESIREN_FG", 256, 32, 4, 256)
    meta = dict(META, clamp_mode=None)
    with torch.no_grad(), pytest.raises(TypeError):
        gen(make_z(), torch.eye(4).unsqueeze(0), **meta)
    with pytest.raises(TypeError):
        vr.fancy_integration(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 1), "cpu", clamp_mode="tanh")


def test_z_must_be_volume_and_global():
    gen = ImplicitGenerator3d("DOUBLESIREN_FG", 256, 32, 4, 256)
    with torch.no_grad(), pytest.raises(ValueError):
        gen(torch.zeros(1, 32, 8, 8, 8), torch.eye(4).unsqueeze(0), **META)


def test_no_cpu_fallback():
    """CPU tensors must be refused, not rendered by some eager path."""
    gen = ImplicitGenerator3d("DOUBLESIREN_FG", 256, 32, 4, 256)
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU implementation"):
        gen(make_z(), torch.eye(4).unsqueeze(0), **META)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        ops.sample_pdf(torch.zeros(2, 5), torch.ones(2, 4), torch.rand(2, 3))
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        vr.fancy_integration(torch.zeros(1, 2, 4, 4), torch.zeros(1, 2, 4, 1), "cpu", clamp_mode="relu")


def test_camera_tables_bitwise_equal_oracle():
    for img, S in ((8, 6), (64, 12), (128, 24)):
        rays, t = vr.camera_tables((img, img), S, META["fov"], 0.25, 1.95, "cpu")
        _, t_ref, d_ref = oracle.camera_rays(1, S, img, META["fov"], 0.25, 1.95)
        assert torch.equal(rays, d_ref[0]) and torch.equal(t, t_ref[0, 0, :, 0])
    pts, z, d = vr.get_initial_rays_trig(2, 6, "cpu", META["fov"], (8, 8), 0.25, 1.95)
    p_ref, z_ref, d_ref = oracle.camera_rays(2, 6, 8, META["fov"], 0.25, 1.95)
    assert torch.equal(pts, p_ref) and torch.equal(z, z_ref) and torch.equal(d, d_ref)


def test_camera_helpers_match_oracle():
    import numpy as np
    np.random.seed(5)
    o = vr.sample_camera_positions("cpu", "y", 0.7, 1.5, 6)
    o_ref = oracle.random_camera_origins(6, 0.7, 1.5, "y", np.random.RandomState(5))
    assert torch.equal(o, o_ref)
    assert torch.equal(vr.create_cam2world_matrix(o, "y"), oracle.look_at_cam2world(o, "y"))
    assert torch.equal(vr.create_cam2world_matrix(o, "z"), oracle.look_at_cam2world(o, "z"))


def test_fp16_only_classes_reject_bf16_operands():
    """The frequency_init(12) classes are offered with fp16 operands only (bf16 misses the 1e-2 contract)."""
    gen = ImplicitGenerator3d("SHORTSIREN_FG", 256, 32, 4, 256)
    assert gen.siren.precision == "fp16"
    with pytest.raises(ValueError):
        gen.siren.precision = "bf16"
    gen.siren.precision = "fp32"
    with pytest.raises(ValueError):
        gen.siren.precision = "fp64"
    tall = ImplicitGenerator3d("TALLSIREN_FG", 256, 32, 4, 256)
    assert tall.siren.precision == "bf16"
    tall.siren.precision = "fp16"


def test_dropout_networks_load_and_only_training_mode_is_refused():
    """siren.py:146-160: nn.Dropout acts in training mode only; a network built with drop_out > 0 keeps the reference's module tree
    (no extra state-dict keys), renders in eval mode, and refuses training mode loudly (the mask is not built into the kernels)."""
    gen = ImplicitGenerator3d("TALLSIREN_FG", 256, 32, 4, 256, drop_out=0.1)
    plain = ImplicitGenerator3d("TALLSIREN_FG", 256, 32, 4, 256)
    assert set(gen.state_dict()) == set(plain.state_dict())
    gen.eval()
    gen.siren.check_dropout()
    gen.train()
    with pytest.raises(NotImplementedError):
        gen.siren.check_dropout()
    plain.train()
    plain.siren.check_dropout()


def test_alias_classes():
    assert siren.TALLSIREN_dg is siren.TALLSIREN_FG and siren.DoubleSIREN_dg is siren.DOUBLESIREN_FG


def test_unmodulated_variant_layout():
    """SHORTSIREN_F (generators/siren.py:830-904): no mapping network, z is the feature volume alone."""
    gen = ImplicitGenerator3d("SHORTSIREN_F", 256, 32, 4, 256)
    ref_state = oracle.init_generator_state("SHORTSIREN_F")
    assert set(gen.state_dict().keys()) == set(ref_state.keys())
    assert not any("mapping_network" in k for k in ref_state)
    gen.load_state_dict(ref_state, strict=True)
    freq, phase = gen.siren.film_parameters(None, 3, "cpu")
    assert freq.shape == (3, 4 * 256) and bool((freq == 1).all()) and bool((phase == 0).all())
    with torch.no_grad(), pytest.raises(ValueError):
        gen((torch.zeros(1, 32, 8, 8, 8), torch.zeros(1, 256)), torch.eye(4).unsqueeze(0), **META)


def test_residual_variant_layout():
    """TALLSIREN_dRes (generators/siren.py:333-408): the reference's module tree and keys, input_dim = z_dim, no mapping network,
    raw head."""
    gen = ImplicitGenerator3d("TALLSIREN_dRes", 32, 3, 4, 256)          # input_dim is overridden by z_dim, as in the reference
    ref_state = oracle.init_generator_state("TALLSIREN_dRes", input_dim=32)
    assert list(gen.state_dict().keys()) == list(ref_state.keys())
    assert "siren.network.1.fc2.weight" in ref_state and "siren.network.3.layer.bias" in ref_state
    gen.load_state_dict(ref_state, strict=True)
    assert gen.siren.res_save_mask == 0b000101 and gen.siren.res_add_mask == 0b010100 and not gen.siren.sigmoid_rgb
    ws, bs = gen.siren.layer_parameters()
    assert [tuple(w.shape) for w in ws] == [(256, 32)] + [(256, 256)] * 5 and len(bs) == 6
    for name, zd, L, save, add in (("TALLSIREN_dResLong", 32, 10, 0b0001010101, 0b0101010100), ("SHORTSIREN_FRes", 256, 4, 0b0001, 0b0100)):
        g2 = ImplicitGenerator3d(name, zd, 32, 4, 256)
        st = oracle.init_generator_state(name, input_dim=32)
        assert list(g2.state_dict().keys()) == list(st.keys())
        g2.load_state_dict(st, strict=True)
        assert (g2.siren.res_save_mask, g2.siren.res_add_mask, len(g2.siren.linear_layers())) == (save, add, L)
        assert oracle.SIREN_SPECS[name]["res_save"] == save and oracle.SIREN_SPECS[name]["res_add"] == add


def test_dense_grid_samples_match_oracle():
    """extract_shapes.create_samples mirror: same order, same (quirky, non-integer x / y index) coordinates."""
    from conditioned_nerf_gan_b200 import extract_shapes
    for N, origin, length in ((8, (0, 0, 0), 2.0), (16, (0.1, -0.2, 0.0), 1.2)):
        mine, corner, vs = extract_shapes.create_samples(N, origin, length)
        ref, corner_ref, vs_ref = oracle.dense_grid_samples(N, origin, length)
        assert torch.equal(mine, ref) and vs == vs_ref and (corner == corner_ref).all()
    assert mine.shape == (1, 16 ** 3, 3)


def test_video_camera_path_matches_oracle():
    from conditioned_nerf_gan_b200 import inference
    for up in ("y", "z"):
        cam, fov = inference.video_camera_path(64, 8, 0.9, 1.3, up, "cpu")
        o_ref, fov_ref = oracle.video_camera_origins(64, 8, 0.9, 1.3, up)
        assert torch.equal(cam, oracle.look_at_cam2world(o_ref, up)) and (fov == fov_ref).all()
        assert cam.shape == (64, 4, 4) and fov[0] == 60 and fov[-1] == 30
    with pytest.raises(ValueError):
        inference.video_camera_path(30, 8, 0.9, 1.3)
