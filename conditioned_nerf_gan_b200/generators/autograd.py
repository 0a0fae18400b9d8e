"""Autograd bridge of the rendering path (a14: backward of a4-a8).  Filled in by the backward
kernels; until then a call that needs gradients fails loudly instead of silently detaching."""
from __future__ import annotations


def render_with_grad(gen, volume, global_feature, cam2worlds, img_size, fov, ray_start, ray_end, num_steps,
                     hierarchical_sample, kwargs):
    raise NotImplementedError(
        "ImplicitGenerator3d.forward was called with gradients enabled, but the backward kernels of the rendering "
        "path are not built yet; wrap the call in torch.no_grad() (there is no eager-PyTorch fallback)")


def siren_forward_with_grad(net, points, volume, global_feature):
    raise NotImplementedError(
        "siren.forward was called with gradients enabled, but the backward kernels are not built yet; wrap the "
        "call in torch.no_grad() (there is no eager-PyTorch fallback)")
