"""CPU-only checks of the C-ABI library: it builds, loads, exports every symbol that
include/cng_b200.h declares, validates arguments and refuses to compute without a device."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from conditioned_nerf_gan_b200 import _lib, build
from oracle import nerf_path as oracle

HEADER = os.path.join(ROOT, "include", "cng_b200.h")


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"CNG_API\s+[\w\s\*]+?\b(cng_\w+)\s*\(", src)))


def test_header_symbols_all_exported_and_bound(lib):
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in cng_b200.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature in _lib.SIGNATURES"
    assert sorted(_lib.SIGNATURES) == names


def test_abi_version(lib):
    assert lib.cng_abi_version() == 1


def test_camera_tables_host_match_oracle(lib):
    for (w, s, fov, a, b) in [(8, 6, 49.134342641202636, 0.25, 1.95), (64, 12, 30.0, 0.5, 1.5), (5, 3, 60.0, 0.1, 2.0)]:
        rays = np.zeros((w * w, 3), np.float32)
        t = np.zeros((s,), np.float32)
        code = lib.cng_camera_tables_host(w, w, s, fov, a, b, rays.ctypes.data_as(ctypes.c_void_p), t.ctypes.data_as(ctypes.c_void_p))
        assert code == 0
        _, t_ref, d_ref = oracle.camera_rays(1, s, w, fov, a, b)
        assert np.array_equal(t, t_ref[0, 0, :, 0].numpy())
        assert np.abs(rays - d_ref[0].numpy()).max() <= 1.2e-7     # sqrt/div vs torch.norm: <= 1 ulp


def test_argument_errors_without_device(lib):
    # NULL pointers / bad sizes are rejected before any CUDA call
    assert lib.cng_composite_fwd(None, None, None, 4, 8, 0.0, 0, 0, 0, None, None, None, None) == -1
    assert b"NULL" in lib.cng_last_error()
    assert lib.cng_sample_pdf(None, None, None, 1, 4, 4, 1e-5, None, None, None) == -1
    assert lib.cng_camera_tables_host(0, 4, 4, 30.0, 0.1, 1.0, None, None) == -1
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.cng_composite_fwd(p, p, None, 1, 4, 0.0, 7, 0, 0, p, p, p, None) == -1     # unknown clamp mode
    assert b"clamp mode" in lib.cng_last_error()
    assert lib.cng_composite_fwd(p, p, None, 1, 2000, 0.0, 0, 0, 0, p, p, p, None) == -2  # S > 1024: unsupported
    # empty inputs are a no-op success, even without a device
    assert lib.cng_composite_fwd(p, p, None, 0, 4, 0.0, 0, 0, 0, p, p, p, None) == 0
    assert lib.cng_sample_pdf(p, p, p, 0, 4, 4, 1e-5, p, None, None) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_compute_calls_fail_loudly_without_gpu(lib):
    assert lib.cng_device_check() != 0
    buf = (ctypes.c_float * 64)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    code = lib.cng_composite_fwd(p, p, None, 1, 4, 0.0, 0, 0, 0, p, p, p, None)
    assert code != 0 and lib.cng_last_error() != b""
    with pytest.raises(_lib.CngError):
        _lib.call("cng_sample_pdf", p, p, p, 1, 4, 4, 1e-5, p, None, None)


def test_workspace_query(lib):
    assert lib.cng_film_siren_workspace_bytes(2, 32, 256, 8, _lib.PREC_FP32) == 0
    per_item = (2 + 4 * 7) * 32768 + 8192
    assert lib.cng_film_siren_workspace_bytes(2, 32, 256, 8, _lib.PREC_BF16) == 2 * per_item + 2 * 8 * 256 * 4
    assert lib.cng_film_siren_workspace_bytes(2, 32, 256, 8, _lib.PREC_FP16) == 2 * per_item + 2 * 8 * 256 * 4
