"""B200-native rendering hot path of the conditioned pi-GAN style NeRF-GAN.

    from conditioned_nerf_gan_b200.generators import ImplicitGenerator3d

is a drop-in for ``generators.ImplicitGenerator3d`` of zzhuolun/conditioned-nerf-gan; the kernels
live in ``libcng_b200.so`` (C ABI: include/cng_b200.h, sources: csrc/*.cu, built by ``build.py``).
"""
from . import _lib  # noqa: F401

__version__ = "0.1.0"
