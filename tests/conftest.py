import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Return (arrays-as-torch dict, meta dict or None) for tests/golden/<name>.npz."""
    raw = np.load(os.path.join(GOLDEN, name + ".npz"))
    out, meta = {}, None
    for k in raw.files:
        a = raw[k]
        if a.dtype.kind in "US":
            if k == "meta/json":
                meta = json.loads(str(a))
            else:
                out[k] = json.loads(str(a))
        else:
            out[k] = torch.from_numpy(a)
    return out, meta


FORWARD_FIXTURES = ["fwd_TALLSIREN_FG", "fwd_SHORTSIREN_FG", "fwd_DOUBLESIREN_FG", "fwd_SingleSIREN_dg", "fwd_SHORTSIREN_F", "fwd_TALLSIREN_dRes", "fwd_TALLSIREN_dResLong", "fwd_SHORTSIREN_FRes"]


DENSE_FIXTURES = ["fwd_dense_TALLSIREN_FG", "fwd_dense_SHORTSIREN_FG", "fwd_dense_DOUBLESIREN_FG", "fwd_dense_SingleSIREN_dg"]


def fixture_inputs(name):
    """Rebuild (state, z, cam2world, draws, meta, taps) of a forward fixture."""
    from oracle import nerf_path as oracle

    fx, meta = load_golden(name)
    seed = meta.pop("seed")
    siren_type = meta.pop("siren_type")
    state = oracle.init_generator_state(siren_type, 256, 32, 256, seed=seed)
    gains = meta.pop("dense_head_gains", None)
    if gains is not None:
        state = oracle.dense_head_state(state, *gains)
    checksum = sum(float(v.double().abs().sum()) for v in state.values())
    assert abs(checksum - float(fx["state/checksum"])) < 1e-9 * checksum, "torch CPU generator drifted"
    film = oracle.SIREN_SPECS[oracle.resolve_siren_type(siren_type)].get("film", True)
    z = (fx["in/volume"], fx["in/global"]) if film else fx["in/volume"]      # unmodulated variants take the volume alone
    draws = {k[5:]: v for k, v in fx.items() if k.startswith("draw/")}
    taps = {k[4:]: v for k, v in fx.items() if k.startswith("tap/")}
    return state, siren_type, z, fx["in/cam2world"], draws, meta, taps


def latent_fixture_inputs(name="fwd_SHORTSIREN"):
    """(state, latent z, cam2world, draws, meta, taps) of the position-input SHORTSIREN fixture."""
    from oracle import nerf_path as oracle

    fx, meta = load_golden(name)
    seed, z_dim, gains = meta.pop("seed"), meta.pop("z_dim"), meta.pop("dense_head_gains")
    meta.pop("siren_type")
    state = oracle.dense_head_state(oracle.init_generator_state("SHORTSIREN", z_dim=z_dim, input_dim=3, hidden_dim=256, seed=seed), *gains)
    checksum = sum(float(v.double().abs().sum()) for v in state.values())
    assert abs(checksum - float(fx["state/checksum"])) < 1e-9 * checksum, "torch CPU generator drifted"
    draws = {k[5:]: v for k, v in fx.items() if k.startswith("draw/")}
    taps = {k[4:]: v for k, v in fx.items() if k.startswith("tap/")}
    return state, fx["in/latent"], fx["in/cam2world"], draws, meta, taps


LIBRARY_FIXTURES = ["fwd_TALLSIREN", "fwd_TALLSIREN_dgx", "fwd_SHORTSIREN_FG_Pyrmd"]


def library_fixture_inputs(name):
    """(siren_type, state, z, cam2world, draws, meta, taps, (z_dim, input_dim)) of a library-MLP decoder fixture."""
    from oracle import nerf_path as oracle

    fx, meta = load_golden(name)
    siren_type = meta.pop("siren_type")
    state, z, cam, g = oracle.library_case_inputs(siren_type)
    checksum = sum(float(v.double().abs().sum()) for v in state.values())
    assert abs(checksum - float(fx["state/checksum"])) < 1e-9 * checksum, "torch CPU generator drifted"
    draws = {k[5:]: v for k, v in fx.items() if k.startswith("draw/")}
    taps = {k[4:]: v for k, v in fx.items() if k.startswith("tap/")}
    return siren_type, state, z, cam, draws, meta, taps, oracle.LIBRARY_CASES[siren_type][:2]
