"""3D U-Net voxel encoder: the producer of the rendering path's input.

Mirrors ``generators/unet3d.py:793-827`` (``UNet3D`` = ``Abstract3DUNet`` with ``DoubleConv`` blocks, ``:488-638``)
closely enough that a reference checkpoint loads strictly: same module tree, hence the same state-dict keys
(``encoders.{i}.basic_module.SingleConv{1,2}.{groupnorm,conv}.*``, ``decoders.{i}. ...``, ``final_conv.*``),
same constructor arguments (``configs/thousand/special.py:53-62``), same return convention
(``fv`` or ``(fv, global)``, ``unet3d.py:635-638``).

What is different, B200-first:
  * the convolutions are library calls (cuDNN through ``torch``, as SURVEY.md section 2 row 5 prescribes) but the
    whole network runs in ``channels_last_3d`` memory format, so the final 1x1x1 convolution writes the feature
    volume directly as NDHWC -- the layout the ray-march/gather kernel reads (one trilinear corner of all 32
    channels = one 128-byte line).  ``ImplicitGenerator3d`` takes such a tensor zero-copy (no
    ``cng_volume_to_channels_last`` launch, no 2 x 33.5 MB per image of layout traffic) and its backward hands the
    scatter-added volume gradient back in the same layout;
  * the bottleneck's global feature is one ``mean`` over the spatial axes instead of building an ``nn.AvgPool3d``
    module per call (``unet3d.py:616-619``);
  * GroupNorm (``nn.GroupNorm`` inside every ``SingleConv``, ``unet3d.py:21-132``) runs on the library's own channels-last
    kernels (``cng_group_norm_fwd / _bwd``, csrc/group_norm.cu): ATen's group norm reads a channels-last tensor with a
    stride of D*H*W elements per channel (0.44 ms per call on B200, 13 % of a batch-4 train step) after autocast has
    widened it to fp32; the kernels here read and write the 16-bit tensor as it is, 16 bytes per access.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import torch
import torch.nn as nn
import torch.nn.functional as F


FORCE_CONTIGUOUS = False     # tools/bench_unet.py: run the network NCDHW (the library then converts layouts per convolution)
NATIVE_GROUP_NORM = True     # False: torch's F.group_norm everywhere (A/B and the CPU tests' path)


class _GroupNormCL(torch.autograd.Function):
    """GroupNorm on a channels-last CUDA tensor through cng_group_norm_fwd / cng_group_norm_bwd (output in the input's dtype)."""

    @staticmethod
    def forward(ctx, x, weight, bias, num_groups, eps):
        from .. import ops
        y, mean, rstd = ops.group_norm_channels_last(x, num_groups, weight, bias, eps)
        ctx.save_for_backward(x, weight, mean, rstd)
        ctx.num_groups = num_groups
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        from .. import ops
        x, weight, mean, rstd = ctx.saved_tensors
        fmt = torch.channels_last_3d if x.dim() == 5 else torch.channels_last
        dy = dy.to(x.dtype).contiguous(memory_format=fmt)
        dx, ds, db = ops.group_norm_channels_last_bwd(dy, x, ctx.num_groups, weight, mean, rstd)
        d_w = ds.sum(0).to(weight.dtype) if (weight is not None and ctx.needs_input_grad[1]) else None
        d_b = db.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return dx, d_w, d_b, None, None


class GroupNormCL(nn.GroupNorm):
    """``nn.GroupNorm`` (same parameters, same state-dict keys) that serves channels-last CUDA tensors with the library's kernels;
    anything else (CPU, NCDHW) goes to ``F.group_norm``.  Under autocast the result keeps the input's dtype: the convolution
    that follows would round torch's fp32 result to the same 16-bit values."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        fmt = torch.channels_last_3d if x.dim() == 5 else (torch.channels_last if x.dim() == 4 else None)
        if (NATIVE_GROUP_NORM and x.is_cuda and fmt is not None and x.dtype in (torch.float32, torch.float16, torch.bfloat16)
                and x.is_contiguous(memory_format=fmt) and x.shape[1] <= 1024 and self.num_groups <= 64):
            cpg = x.shape[1] // self.num_groups
            v = 8 if (x.dtype != torch.float32 and cpg % 8 == 0) else (4 if cpg % 4 == 0 else (2 if cpg % 2 == 0 else 1))
            if x.shape[1] // v <= 256:
                with torch.autocast("cuda", enabled=False):
                    return _GroupNormCL.apply(x, self.weight, self.bias, self.num_groups, self.eps)
        return super().forward(x)


def number_of_features_per_level(init_channel_number: int, num_levels: int) -> List[int]:
    """unet3d.py:13-14"""
    return [init_channel_number * 2 ** k for k in range(num_levels)]


class SingleConv(nn.Sequential):
    """One conv layer with its normalisation / non-linearity in the given order (unet3d.py:21-132).
    Letters: c conv3d, r ReLU, l LeakyReLU(0.1), e ELU, g GroupNorm, b BatchNorm3d."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, order: str = "gcr", num_groups: int = 8, padding: int = 1):
        super().__init__()
        if "c" not in order:
            raise AssertionError("Conv layer MUST be present")
        if order[0] in "rle":
            raise AssertionError("Non-linearity cannot be the first operation in the layer")
        for i, ch in enumerate(order):
            before_conv = i < order.index("c")
            if ch == "c":
                bias = not ("g" in order or "b" in order)        # learnable bias only without a normalisation layer
                self.add_module("conv", nn.Conv3d(in_channels, out_channels, kernel_size, padding=padding, bias=bias))
            elif ch == "g":
                channels = in_channels if before_conv else out_channels
                groups = 1 if channels < num_groups else num_groups
                if channels % groups:
                    raise AssertionError(f"Expected number of channels in input to be divisible by num_groups. num_channels={channels}, num_groups={groups}")
                self.add_module("groupnorm", GroupNormCL(num_groups=groups, num_channels=channels))
            elif ch == "b":
                self.add_module("batchnorm", nn.BatchNorm3d(in_channels if before_conv else out_channels))
            elif ch == "r":
                self.add_module("ReLU", nn.ReLU(inplace=True))
            elif ch == "l":
                self.add_module("LeakyReLU", nn.LeakyReLU(negative_slope=0.1, inplace=True))
            elif ch == "e":
                self.add_module("ELU", nn.ELU(inplace=True))
            else:
                raise ValueError(f"Unsupported layer type '{ch}'. MUST be one of ['b', 'g', 'r', 'l', 'e', 'c']")


class DoubleConv(nn.Sequential):
    """Two SingleConv layers; channel plan of unet3d.py:157-192."""

    def __init__(self, in_channels: int, out_channels: int, encoder: bool, kernel_size: int = 3, order: str = "gcr", num_groups: int = 8):
        super().__init__()
        if encoder:
            mid = max(out_channels // 2, in_channels)
            plan = ((in_channels, mid), (mid, out_channels))
        else:
            plan = ((in_channels, out_channels), (out_channels, out_channels))
        for i, (ci, co) in enumerate(plan, start=1):
            self.add_module(f"SingleConv{i}", SingleConv(ci, co, kernel_size, order, num_groups))


class ExtResNetBlock(nn.Module):
    """SingleConv followed by a two-convolution residual block, the non-linearity of the last one applied after the sum
    (unet3d.py:195-265)."""

    def __init__(self, in_channels: int, out_channels: int, kernel_size: int = 3, order: str = "cge", num_groups: int = 8, **kwargs):
        super().__init__()
        self.conv1 = SingleConv(in_channels, out_channels, kernel_size=kernel_size, order=order, num_groups=num_groups)
        self.conv2 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=order, num_groups=num_groups)
        n_order = order
        for c in "rel":
            n_order = n_order.replace(c, "")
        self.conv3 = SingleConv(out_channels, out_channels, kernel_size=kernel_size, order=n_order, num_groups=num_groups)
        if "l" in order:
            self.non_linearity = nn.LeakyReLU(negative_slope=0.1, inplace=True)
        elif "e" in order:
            self.non_linearity = nn.ELU(inplace=True)
        else:
            self.non_linearity = nn.ReLU(inplace=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        out = self.conv1(x)
        residual = out
        out = self.conv3(self.conv2(out))
        return self.non_linearity(out + residual)


class Encoder(nn.Module):
    """Optional 2x2x2 max pooling + basic module (unet3d.py:268-323)."""

    def __init__(self, in_channels: int, out_channels: int, apply_pooling: bool = True, basic_module=DoubleConv,
                 conv_layer_order: str = "gcr", num_groups: int = 8):
        super().__init__()
        self.pooling = nn.MaxPool3d(kernel_size=(2, 2, 2)) if apply_pooling else None
        self.basic_module = basic_module(in_channels, out_channels, encoder=True, kernel_size=3, order=conv_layer_order, num_groups=num_groups)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if self.pooling is not None:
            x = self.pooling(x)
        return self.basic_module(x)


class _TransposedUpsampling(nn.Module):
    """Parameter holder + forward of the learned upsampling (unet3d.py:405-452): ConvTranspose3d(kernel 3, stride 2,
    padding 1) to the skip connection's size.  Key: ``upsample.{weight,bias}``."""

    def __init__(self, in_channels: int, out_channels: int):
        super().__init__()
        self.upsample = nn.ConvTranspose3d(in_channels, out_channels, kernel_size=3, stride=(2, 2, 2), padding=1)

    def forward(self, encoder_features: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        return self.upsample(x, encoder_features.size()[2:])


class Decoder(nn.Module):
    """DoubleConv: nearest-neighbour upsampling to the skip connection's size, channel concatenation (skip first);
    ExtResNetBlock: transposed-convolution upsampling, summation (unet3d.py:326-403, 405-452)."""

    def __init__(self, in_channels: int, out_channels: int, basic_module=DoubleConv, conv_layer_order: str = "gcr", num_groups: int = 8):
        super().__init__()
        self.concat = basic_module is DoubleConv
        if not self.concat:
            self.upsampling = _TransposedUpsampling(in_channels, out_channels)
            in_channels = out_channels
        self.basic_module = basic_module(in_channels, out_channels, encoder=False, kernel_size=3, order=conv_layer_order, num_groups=num_groups)

    def forward(self, encoder_features: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
        if self.concat:
            x = F.interpolate(x, size=encoder_features.shape[2:], mode="nearest")
            return self.basic_module(torch.cat((encoder_features, x), dim=1))
        return self.basic_module(encoder_features + self.upsampling(encoder_features, x))


class _UNet3DBase(nn.Module):
    """``UNet3D(in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr", num_groups=8, num_levels=4,
    is_segmentation=True, return_global=False)`` -- unet3d.py:793-827.

    ``forward(voxels[B, in_channels, D, H, W])`` returns the feature volume ``[B, out_channels, D, H, W]`` (logical NCDHW
    shape, NDHWC strides) or ``(volume, global[B, f_maps[-1]])`` when ``return_global``."""

    def __init__(self, in_channels: int, out_channels: int, final_sigmoid: bool = True, f_maps: Union[int, Sequence[int]] = 64,
                 layer_order: str = "gcr", num_groups: int = 8, num_levels: int = 4, is_segmentation: bool = True,
                 testing: bool = False, return_global: bool = False, basic_module=DoubleConv, pyramid: bool = False, **kwargs):
        super().__init__()
        self.testing = testing
        self.pyramid = pyramid
        if isinstance(f_maps, int):
            f_maps = number_of_features_per_level(f_maps, num_levels=num_levels)
        f_maps = list(f_maps)
        self.encoders = nn.ModuleList(
            Encoder(in_channels if i == 0 else f_maps[i - 1], f, apply_pooling=i > 0, basic_module=basic_module,
                    conv_layer_order=layer_order, num_groups=num_groups)
            for i, f in enumerate(f_maps))
        rev = f_maps[::-1]
        self.decoders = nn.ModuleList(
            Decoder(rev[i] + rev[i + 1] if basic_module is DoubleConv else rev[i], rev[i + 1], basic_module=basic_module,
                    conv_layer_order=layer_order, num_groups=num_groups) for i in range(len(rev) - 1))
        if not pyramid:                                    # the pyramid network has no final convolution (unet3d.py:640-791)
            self.final_conv = nn.Conv3d(f_maps[0], out_channels, 1)
            if is_segmentation:
                self.final_activation = nn.Sigmoid() if final_sigmoid else nn.Softmax(dim=1)
            else:
                self.final_activation = None
        self.return_global = return_global

    def forward(self, x: torch.Tensor):
        if x.is_cuda and not FORCE_CONTIGUOUS:       # cuDNN runs the whole network NDHWC; the CPU kernels (tests only) stay NCDHW until the output
            x = x.contiguous(memory_format=torch.channels_last_3d)
        skips = []
        for encoder in self.encoders:
            x = encoder(x)
            skips.insert(0, x)
        global_features = x.mean(dim=(2, 3, 4)) if self.return_global else None      # full-extent average pool of the bottleneck
        if self.pyramid:
            levels = []
            for decoder, skip in zip(self.decoders, skips[1:]):
                x = decoder(skip, x)
                levels.append(x)
            return (levels, global_features) if self.return_global else levels
        for decoder, skip in zip(self.decoders, skips[1:]):
            x = decoder(skip, x)
        x = self.final_conv(x)
        if self.testing and self.final_activation is not None:
            x = self.final_activation(x)
        if x.dim() == 5 and not x.is_contiguous(memory_format=torch.channels_last_3d):
            x = x.contiguous(memory_format=torch.channels_last_3d)                    # the layout the gather kernel reads
        return (x, global_features) if self.return_global else x


class UNet3D(_UNet3DBase):
    """unet3d.py:793-827: DoubleConv blocks, nearest-neighbour upsampling, concatenation joining."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr", num_groups=8, num_levels=4,
                 is_segmentation=True, return_global=False, **kwargs):
        super().__init__(in_channels, out_channels, final_sigmoid=final_sigmoid, f_maps=f_maps, layer_order=layer_order, num_groups=num_groups,
                         num_levels=num_levels, is_segmentation=is_segmentation, return_global=return_global, basic_module=DoubleConv, **kwargs)


class PyramidUNet3D(_UNet3DBase):
    """unet3d.py:829-863: the same network without the final convolution; returns every decoder level (coarse to fine)."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr", num_groups=8, num_levels=4,
                 is_segmentation=True, return_global=False, **kwargs):
        super().__init__(in_channels, out_channels, final_sigmoid=final_sigmoid, f_maps=f_maps, layer_order=layer_order, num_groups=num_groups,
                         num_levels=num_levels, is_segmentation=is_segmentation, return_global=return_global, basic_module=DoubleConv,
                         pyramid=True, **kwargs)


class ResidualUNet3D(_UNet3DBase):
    """unet3d.py:865-898: ExtResNetBlock blocks, transposed-convolution upsampling, summation joining, five levels by default."""

    def __init__(self, in_channels, out_channels, final_sigmoid=True, f_maps=64, layer_order="gcr", num_groups=8, num_levels=5,
                 is_segmentation=True, return_global=False, **kwargs):
        super().__init__(in_channels, out_channels, final_sigmoid=final_sigmoid, f_maps=f_maps, layer_order=layer_order, num_groups=num_groups,
                         num_levels=num_levels, is_segmentation=is_segmentation, return_global=return_global, basic_module=ExtResNetBlock, **kwargs)
