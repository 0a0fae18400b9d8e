"""ctypes binding of libcng_b200.so (include/cng_b200.h).  No torch types cross this boundary.

The library is built in-tree by ``conditioned_nerf_gan_b200.build`` (nvcc, sm_100a).  Loading
fails loudly if it is missing: there is no Python / CPU fallback for any compute entry point.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_longlong, c_size_t, c_void_p

_PKG = os.path.dirname(os.path.abspath(__file__))
# CNG_LIB points at an alternative build of the same ABI (A/B experiments of kernel variants in one gpurun call)
LIB_PATH = os.environ.get("CNG_LIB") or os.path.join(_PKG, "libcng_b200.so")

CNG_OK = 0
CLAMP_RELU, CLAMP_SOFTPLUS = 0, 1
PREC_FP32, PREC_BF16, PREC_FP16 = 0, 1, 2

# name -> (restype, argtypes); must list every symbol declared in include/cng_b200.h
SIGNATURES = {
    "cng_abi_version": (c_int, []),
    "cng_last_error": (c_char_p, []),
    "cng_device_check": (c_int, []),
    "cng_camera_tables_host": (c_int, [c_int, c_int, c_int, c_double, c_double, c_double, c_void_p, c_void_p]),
    "cng_volume_to_channels_last": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cng_volume_f16_to_channels_last": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cng_raymarch_gather_coarse": (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                           c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cng_raymarch_gather_fine": (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cng_gather_points": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p]),
    "cng_film_parameters": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cng_film_siren_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "cng_film_siren_fwd": (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "cng_film_siren_fwd_gather": (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]),
    "cng_film_siren_res_scratch_bytes": (c_size_t, []),
    "cng_film_siren_fwd_res": (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int, c_int, ctypes.c_uint, ctypes.c_uint, c_void_p, c_size_t, c_void_p, c_size_t,
                                       c_void_p, c_void_p]),
    "cng_composite_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_int, c_int, c_int,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "cng_sample_pdf": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "cng_resample_from_coarse": (c_int, [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    "cng_merge_composite_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                        c_float, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cng_scatter_points": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_longlong, c_void_p, c_void_p]),
    "cng_volume_from_channels_last": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "cng_film_siren_fwd_train": (c_int, [c_void_p, c_int, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_int, c_int, ctypes.c_uint, ctypes.c_uint, c_void_p, c_size_t, c_void_p, c_size_t,
                                         c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cng_film_siren_wt_image_bytes": (c_size_t, [c_int]),
    "cng_film_siren_g_dump_bits": (c_int, []),
    "cng_film_siren_wt_images": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "cng_film_siren_dgrad": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_void_p, c_void_p, c_longlong, c_void_p, c_void_p, c_void_p,
                                     ctypes.c_uint, ctypes.c_uint, c_void_p, c_size_t, c_void_p]),
    "cng_film_siren_wgrad": (c_int, [c_void_p, c_void_p, c_longlong, c_void_p, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cng_film_siren_head_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_longlong, c_int, c_void_p, c_void_p]),
    "cng_film_siren_bwd_workspace_bytes": (c_size_t, [c_longlong, c_int, c_int, c_int]),
    "cng_film_siren_bwd": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p, c_int, ctypes.c_uint, ctypes.c_uint, c_void_p, c_size_t, c_void_p,
                                   c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cng_group_norm_fwd": (c_int, [c_void_p, c_int, c_int, c_longlong, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "cng_group_norm_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_longlong, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_void_p, c_void_p]),
    "cng_composite_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_int, c_int, c_int,
                                  c_void_p, c_void_p]),
    "cng_render_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "cng_render_fwd": (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                               c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_int, c_float, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "cng_merge_sort": (c_int, [c_void_p, c_void_p, c_longlong, c_int, c_void_p, c_void_p, c_void_p]),
    "cng_merge_composite": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float,
                                    c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


class CngError(RuntimeError):
    """A libcng_b200 entry point returned a non-zero status."""

    def __init__(self, fn: str, code: int, message: str):
        super().__init__(f"{fn} failed with status {code}: {message}")
        self.fn, self.code, self.message = fn, code, message


def load() -> ctypes.CDLL:
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m conditioned_nerf_gan_b200.build` "
            "(nvcc, sm_100a). There is no CPU or PyTorch fallback for the rendering path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == ABI mismatch with the header
        fn.restype, fn.argtypes = res, args
    if lib.cng_abi_version() != 1:
        raise RuntimeError(f"libcng_b200 ABI version {lib.cng_abi_version()} != 1")
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point and raise CngError on failure."""
    lib = load()
    code = getattr(lib, name)(*args)
    if code != CNG_OK:
        raise CngError(name, code, lib.cng_last_error().decode(errors="replace"))
