"""Pin the oracle (oracle/nerf_path.py) against golden vectors recorded from the unmodified
reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import DENSE_FIXTURES, FORWARD_FIXTURES, LIBRARY_FIXTURES, fixture_inputs, latent_fixture_inputs, library_fixture_inputs, load_golden
from oracle import nerf_path as oracle


@pytest.mark.parametrize("name", FORWARD_FIXTURES + DENSE_FIXTURES)
def test_forward_matches_reference(name):
    state, siren_type, z, cam, draws, meta, taps = fixture_inputs(name)
    out = oracle.render(state, siren_type, z, cam, draws, **meta)
    # everything except the resampling uses the reference's own torch ops -> bit-identical
    exact = ["points_coarse", "rgb_sigma_coarse"]
    if meta["hierarchical_sample"]:
        exact += ["weights_coarse"]
    for k in exact:
        assert torch.equal(out[k], taps[k].reshape(out[k].shape)), k
    dense = name in DENSE_FIXTURES
    if meta["hierarchical_sample"]:
        # sum(weights) association order differs (see oracle docstring): 1 ulp on the cdf, which
        # (u - cdf_lo) / denom amplifies by up to bin_width / denom (denom >= 1e-5): almost every
        # sample agrees to a few ulp, a rare one in a near-empty bin to ~1e-5.  The dense fixtures have truly empty bins
        # (weight = the 2e-5 floor, denom ~ 1e-5 .. 1e-4): there the bound bin_width * ulp / denom reaches ~1e-3 for a
        # handful of samples -- which carry no weight in the image (pixels agree to 1e-4 below).
        dt = (out["t_fine"] - taps["t_fine"]).abs()
        assert float(dt.max()) < (2e-3 if dense else 1e-4) and float((dt > 2e-6).float().mean()) < (1e-2 if dense else 2e-3)
        assert torch.allclose(out["points_fine"], taps["points_fine"], rtol=0, atol=2e-3 if dense else 1e-4)
        d_fine = (out["rgb_sigma_fine"] - taps["rgb_sigma_fine"]).abs()
        if dense:
            assert float((d_fine > 1e-3).float().mean()) < 0.1      # sigma carries a gain of 300: the moved samples move it
        else:
            assert float(d_fine.max()) < 2e-4
    for k, tol in (("rgb", 1e-5), ("dist", 1e-5), ("pixels", 2e-5), ("depth", 1e-5)):
        assert torch.allclose(out[k], taps[k].reshape(out[k].shape), rtol=0, atol=1e-4 if dense else tol), k
    assert out["pixels"].shape == (cam.shape[0], 3, meta["img_size"], meta["img_size"])
    assert out["depth"].shape == (cam.shape[0], meta["img_size"], meta["img_size"])


def test_latent_shortsiren_matches_reference():
    """SHORTSIREN (position input, latent z through CustomMappingNetwork; siren.py:1172-1224) against the fixture recorded from the
    reference's own class."""
    state, latent, cam, draws, meta, taps = latent_fixture_inputs()
    out = oracle.render(state, "SHORTSIREN", latent, cam, draws, **meta)
    assert torch.equal(out["points_coarse"], taps["points_coarse"]) and torch.equal(out["rgb_sigma_coarse"], taps["rgb_sigma_coarse"])
    assert torch.allclose(out["pixels"], taps["pixels"], rtol=0, atol=1e-4) and torch.allclose(out["depth"], taps["depth"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("name", LIBRARY_FIXTURES)
def test_library_mlp_decoders_match_reference(name):
    """TALLSIREN (per-point FiLM), TALLSIREN_dgx, SHORTSIREN_FG_Pyrmd: the oracle against fixtures recorded from the reference classes."""
    siren_type, state, z, cam, draws, meta, taps, _ = library_fixture_inputs(name)
    out = oracle.render(state, siren_type, z, cam, draws, **meta)
    assert torch.equal(out["points_coarse"], taps["points_coarse"])
    assert torch.allclose(out["rgb_sigma_coarse"], taps["rgb_sigma_coarse"], rtol=0, atol=1e-6)
    assert torch.allclose(out["pixels"], taps["pixels"], rtol=0, atol=1e-4) and torch.allclose(out["depth"], taps["depth"], rtol=0, atol=1e-4)


def test_composite_matches_reference():
    fx, _ = load_golden("functions")
    for i in range(5):
        cfg = fx[f"comp/case{i}/cfg"]
        rgb, dist, w = oracle.composite(fx["comp/rgb_sigma"], fx["comp/t"], fx["comp/noise"], cfg["noise_std"],
                                        cfg["clamp_mode"], cfg["white_back"], cfg["last_back"])
        assert torch.equal(rgb, fx[f"comp/case{i}/rgb"])
        assert torch.equal(dist, fx[f"comp/case{i}/dist"])
        assert torch.equal(w, fx[f"comp/case{i}/weights"])


def test_composite_rejects_unknown_clamp_mode():
    fx, _ = load_golden("functions")
    with pytest.raises(TypeError):
        oracle.composite(fx["comp/rgb_sigma"], fx["comp/t"], fx["comp/noise"], 0.0, None)


def test_resample_pdf_indices_bit_exact_vs_reference():
    fx, _ = load_golden("functions")
    samples, inds, below, above = oracle.resample_pdf(fx["pdf/bins"], fx["pdf/weights"], fx["pdf/u"])
    assert inds.dtype == torch.int64
    assert torch.equal(inds, fx["pdf/inds"]), "bin indices differ from the reference on the golden vector"
    ds = (samples - fx["pdf/samples"]).abs()      # 1-ulp cdf differences amplified by 1/denom
    assert float(ds.max()) < 1e-4 and float((ds > 2e-6).float().mean()) < 2e-3
    M = fx["pdf/weights"].shape[1]
    assert int(inds.min()) >= 0 and int(inds.max()) <= M
    assert torch.equal(below, (inds - 1).clamp_min(0)) and torch.equal(above, inds.clamp_max(M))
    # degenerate rows: all-zero weights give the uniform pdf; the spike row stays inside its bin
    b = fx["pdf/bins"]
    assert (samples[0] >= b[0, 0]).all() and (samples[0] <= b[0, -1]).all()
    assert (samples[1] >= b[1, 5] - 1e-6).all() and (samples[1] <= b[1, 6] + 1e-6).all()


def test_searchsorted_side_example():
    # SURVEY.md appendix A.8 probe
    cdf = torch.tensor([[0, 2.5e-6, 0.25, 0.25001, 1.0]])
    u = torch.tensor([[0, 0.1, 0.25, 0.5, 0.999999, 1.0]])
    assert torch.searchsorted(cdf, u).tolist() == [[0, 2, 2, 4, 4, 4]]


def test_trilinear_manual_matches_grid_sample():
    fx, _ = load_golden("functions")
    vol, pts = fx["tri/volume"], fx["tri/points"]
    feat, idx = oracle.trilinear_manual(vol[0].numpy(), pts[0].numpy())
    ref = fx["tri/features"][0].numpy()
    assert np.abs(feat - ref).max() <= 2e-6
    # and the oracle's grid_sample wrapper is the reference op itself
    again = oracle.trilinear_lookup(vol, pts, 16, 4)
    assert torch.equal(again, fx["tri/features"])
    D, H, W = vol.shape[2:]
    assert idx[:, 0].max() <= W - 1 and idx[:, 1].max() <= H - 1 and idx[:, 2].max() <= D - 1 and idx.min() >= 0


def test_state_dict_keys_and_aliases():
    st = oracle.init_generator_state("TALLSIREN_dg")
    assert st["siren.network.0.layer.weight"].shape == (256, 32)
    assert st["siren.network.7.layer.weight"].shape == (256, 256)
    assert st["siren.final_layer.weight"].shape == (4, 256)
    assert st["siren.mapping_network.weight"].shape == (2 * 8 * 256, 256)
    assert oracle.resolve_siren_type("DoubleSIREN_dg") == "DOUBLESIREN_FG"
    with pytest.raises(AttributeError):
        oracle.resolve_siren_type("NOPE")


def test_draw_order_matches_reference_stream():
    g1 = torch.Generator().manual_seed(3)
    d = oracle.draw_randoms(2, 4, 6, True, g1)
    g2 = torch.Generator().manual_seed(3)
    a = torch.rand((2, 16, 6, 1), generator=g2)
    b = torch.randn((2, 16, 6, 1), generator=g2)
    c = torch.rand((32, 6), generator=g2)
    e = torch.randn((2, 16, 12, 1), generator=g2)
    assert torch.equal(d["u_jitter"], a) and torch.equal(d["noise_coarse"], b)
    assert torch.equal(d["u_resample"], c) and torch.equal(d["noise_final"], e)
