"""Progressive CoordConv discriminator: the consumer of the rendering path's images in training.

Mirrors ``discriminators/discriminators.py:87-199`` (``ProgressiveDiscriminator``) with the same module tree, so the
state-dict keys are the reference's (``layers.{i}.network.{0,2}.conv.*``, ``layers.{i}.proj.*``,
``fromRGB.{i}.model.0.*``, ``final_layer.*``) and checkpoints load strictly; ``forward(input, alpha, instance_noise=0,
cond=None, **kwargs)`` has the reference's signature (every curriculum key is passed and ignored, ``utils.py:663``).

B200-first differences (results agree to fp32 summation order):
  * ``CoordConv`` (``:87-103``) does not materialise ``cat([x, xx, yy])``.  The convolution is linear in its input
    channels, so conv(cat[x, coords]) = conv_x(x) + conv_c(coords): the second term does not depend on the image, is
    computed once per call on a cached ``[1, 2, H, W]`` grid (the reference rebuilds the grid on the CPU and copies it
    to the device in every CoordConv of every forward, ``:49-76``) and is broadcast over the batch.  Zero padding
    commutes with the split, double backward (R1 penalty, ``utils.py:805-813``) goes through stock conv2d;
  * convolutions are cuDNN library calls (out of the hot-path scope), but their autograd is spelled
    out (``conv2d_r1``): the R1 penalty differentiates the image gradient once more, and PyTorch's generic
    ``_convolution_double_backward`` forms the weight term as a convolution with batch and channels swapped -- an
    ``[C, N, H, W] * [C', N, H, W]`` problem with a 128 x 128 "filter" that cuDNN only serves with its legacy SGEMM
    kernels (measured on B200, batch 8 at 128^2: 2 launches of 88 ms, 250 of the 254 ms of one discriminator update).
    Here the input gradient is its own autograd node whose backward is a plain convolution (w.r.t. the upstream
    gradient) and a plain weight-gradient (w.r.t. the filter): the same numbers from tensor-core kernels.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


_R1_INPUT_ONLY = False


class r1_pass:
    """Context manager for the backward pass that forms the R1 penalty's image gradient (``torch.autograd.grad(..., inputs=real
    images, create_graph=True)``, utils.py:805-813): only d / d input is consumed there, so the convolutions skip their weight
    and bias gradients (``needs_input_grad`` is fixed at forward time and would have every layer compute and discard them)."""

    def __enter__(self):
        global _R1_INPUT_ONLY
        self.prev, _R1_INPUT_ONLY = _R1_INPUT_ONLY, True
        return self

    def __exit__(self, *exc):
        global _R1_INPUT_ONLY
        _R1_INPUT_ONLY = self.prev
        return False


def _cl(t):
    """channels_last for 4-D CUDA tensors: the library's tensor-core convolution kernels are NHWC; an NCHW operand costs a
    layout kernel before and after every convolution (measured: 550 of them, 6 % of a batch-4 train step)."""
    return t.contiguous(memory_format=torch.channels_last) if (t.is_cuda and t.dim() == 4) else t.contiguous()


class _ConvInputGrad(torch.autograd.Function):
    """gx = d conv2d(x, w) / dx contracted with gy (a transposed convolution), differentiable a second time:
    d gx / d gy contracted with ggx = conv2d(ggx, w);  d gx / d w contracted with ggx = weight-gradient(ggx, gy)."""

    @staticmethod
    def forward(ctx, gy, w, x_shape, stride, padding):
        ctx.save_for_backward(gy, w)
        ctx.cfg = (x_shape, stride, padding)
        return torch.nn.grad.conv2d_input(x_shape, w, gy, stride=stride, padding=padding)

    @staticmethod
    def backward(ctx, ggx):
        gy, w = ctx.saved_tensors
        _, stride, padding = ctx.cfg
        ggx = _cl(ggx.to(w.dtype))
        d_gy = F.conv2d(ggx, w, None, stride, padding) if ctx.needs_input_grad[0] else None
        d_w = torch.nn.grad.conv2d_weight(ggx, w.shape, gy, stride=stride, padding=padding) if ctx.needs_input_grad[1] else None
        return d_gy, d_w, None, None, None


class _Conv2dR1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, stride, padding):
        ctx.save_for_backward(x, w)
        ctx.cfg = (stride, padding, b is not None)
        return F.conv2d(x, w, b, stride, padding)

    @staticmethod
    def backward(ctx, gy):
        x, w = ctx.saved_tensors
        stride, padding, has_bias = ctx.cfg
        gy = _cl(gy)
        params = not _R1_INPUT_ONLY
        gx = _ConvInputGrad.apply(gy, w, x.shape, stride, padding) if ctx.needs_input_grad[0] else None
        gw = torch.nn.grad.conv2d_weight(x, w.shape, gy, stride=stride, padding=padding) if (params and ctx.needs_input_grad[1]) else None
        gb = gy.sum((0, 2, 3)) if (params and has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None, None


def conv2d_r1(x, w, b=None, stride=1, padding=0):
    """``F.conv2d`` (groups 1, dilation 1) whose double backward stays on the library's fast kernels (module docstring).
    Follows autocast: operands are cast once here, the function itself runs with autocast off."""
    stride = (stride, stride) if isinstance(stride, int) else tuple(stride)
    padding = (padding, padding) if isinstance(padding, int) else tuple(padding)
    if x.is_cuda and torch.is_autocast_enabled():
        dt = torch.get_autocast_dtype('cuda')
        x, w, b = _cl(x.to(dt)), _cl(w.to(dt)), (b.to(dt) if b is not None else None)
        with torch.autocast("cuda", enabled=False):
            return _Conv2dR1.apply(x, w, b, stride, padding)
    if w.dtype != x.dtype:
        w, b = w.to(x.dtype), (b.to(x.dtype) if b is not None else None)
    return _Conv2dR1.apply(_cl(x), _cl(w), b, stride, padding)


class R1Conv2d(nn.Conv2d):
    """``nn.Conv2d`` (same parameters, same state-dict keys) routed through ``conv2d_r1``."""

    def forward(self, x):
        return conv2d_r1(x, self.weight, self.bias, self.stride, self.padding)


class GlobalAveragePooling(nn.Module):
    def forward(self, x):
        return x.mean([2, 3])


class AdapterBlock(nn.Module):
    """1x1 conv + LeakyReLU(0.2) from image channels (discriminators.py:21-30)."""

    def __init__(self, output_channels: int, input_channels: int = 3):
        super().__init__()
        self.model = nn.Sequential(R1Conv2d(input_channels, output_channels, 1, padding=0), nn.LeakyReLU(0.2))

    def forward(self, input):
        return self.model(input)


def kaiming_leaky_init(m):
    """discriminators.py:32-37 (matches Linear layers only -- as in the reference, the conv layers keep their default init)."""
    if m.__class__.__name__.find("Linear") != -1:
        torch.nn.init.kaiming_normal_(m.weight, a=0.2, mode="fan_in", nonlinearity="leaky_relu")


_GRID_CACHE: Dict[Tuple, torch.Tensor] = {}


def coord_grid(h: int, w: int, device, dtype) -> torch.Tensor:
    """[1, 2, h, w]: channel 0 varies along dim 2 (rows) as -1 + 2 i / (h - 1), channel 1 along dim 3 -- the two
    channels ``AddCoords`` appends (discriminators.py:49-66; its x_dim is the tensor's dim 2)."""
    key = (h, w, str(device), dtype)
    g = _GRID_CACHE.get(key)
    if g is None:
        ii = torch.arange(h, device=device, dtype=torch.float32) / (h - 1) * 2 - 1
        jj = torch.arange(w, device=device, dtype=torch.float32) / (w - 1) * 2 - 1
        g = torch.stack([ii[:, None].expand(h, w), jj[None, :].expand(h, w)])[None].to(dtype).contiguous()
        _GRID_CACHE[key] = g
    return g


class CoordConv(nn.Module):
    """conv2d over [x, xx, yy] without building the concatenation (see the module docstring)."""

    def __init__(self, in_channels: int, out_channels: int, with_r: bool = False, **kwargs):
        super().__init__()
        if with_r:
            raise NotImplementedError("with_r is never used by the reference's discriminators")
        self.in_channels = in_channels
        self.conv = nn.Conv2d(in_channels + 2, out_channels, **kwargs)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        c = self.conv
        w = c.weight
        y = conv2d_r1(x, w[:, : self.in_channels], None, c.stride, c.padding)
        grid = coord_grid(x.shape[2], x.shape[3], x.device, torch.float32)
        pos = F.conv2d(grid.to(w.dtype), w[:, self.in_channels:], c.bias, c.stride, c.padding)      # no image on this path: stock autograd
        return y + pos.to(y.dtype)


class ResidualCoordConvBlock(nn.Module):
    """discriminators.py:106-135."""

    def __init__(self, inplanes: int, planes: int, kernel_size: int = 3, stride: int = 1, downsample: bool = False, groups: int = 1):
        super().__init__()
        p = kernel_size // 2
        self.network = nn.Sequential(
            CoordConv(inplanes, planes, kernel_size=kernel_size, stride=stride, padding=p),
            nn.LeakyReLU(0.2, inplace=True),
            CoordConv(planes, planes, kernel_size=kernel_size, padding=p),
            nn.LeakyReLU(0.2, inplace=True),
        )
        self.network.apply(kaiming_leaky_init)
        self.proj = R1Conv2d(inplanes, planes, 1) if inplanes != planes else None
        self.downsample = downsample

    def forward(self, identity):
        y = self.network(identity)
        if self.downsample:
            y = F.avg_pool2d(y, 2)
            identity = F.avg_pool2d(identity, 2)
        identity = identity if self.proj is None else self.proj(identity)
        return (y + identity) / math.sqrt(2)


class ProgressiveDiscriminator(nn.Module):
    """discriminators.py:138-199: eight residual CoordConv blocks, entered at the block matching the image size, with the
    progressive-GAN fade-in of the next-lower resolution after the first block."""
    image_channels = 3        # channels the fromRGB adapters read
    final_channels = 1        # outputs of the 2x2 head convolution

    def __init__(self, **kwargs):
        super().__init__()
        self.epoch = 0
        self.step = 0
        planes = [16, 32, 64, 128, 256, 400, 400, 400, 400]
        self.layers = nn.ModuleList(ResidualCoordConvBlock(planes[i], planes[i + 1], downsample=True) for i in range(8))
        self.fromRGB = nn.ModuleList(AdapterBlock(p, self.image_channels) for p in planes)
        self.final_layer = R1Conv2d(400, self.final_channels, 2)
        self.img_size_to_layer = {2: 8, 4: 7, 8: 6, 16: 5, 32: 4, 64: 3, 128: 2, 256: 1, 512: 0}

    def _trunk(self, input, alpha):
        start = self.img_size_to_layer[input.shape[-1]]
        x = self.fromRGB[start](input)
        for i, layer in enumerate(self.layers[start:]):
            if i == 1:
                x = alpha * x + (1 - alpha) * self.fromRGB[start + 1](F.interpolate(input, scale_factor=0.5, mode="nearest"))
            x = layer(x)
        return self.final_layer(x)

    def forward(self, input, alpha, instance_noise=0, cond=None, **kwargs):
        x = self._trunk(input, alpha)
        return x.reshape(x.shape[0], 1)


class ProgressiveEncoderDiscriminator(ProgressiveDiscriminator):
    """discriminators.py:202-271: the same trunk with a 1 + 256 + 2 channel head; also predicts a latent code and a camera
    position, optional instance noise on the input.  Returns (prediction [B,1], latent [B,256], position [B,2])."""
    final_channels = 1 + 256 + 2

    def forward(self, input, alpha, instance_noise=0, cond=None, **kwargs):
        if instance_noise > 0:
            input = input + torch.randn_like(input) * instance_noise
        x = self._trunk(input, alpha)
        x = x.reshape(x.shape[0], -1)
        return x[..., 0:1], x[..., 1:257], x[..., 257:259]


class ProgressiveDiscriminator_inputCat(ProgressiveDiscriminator):
    """discriminators.py:274-335: conditional variant, the condition image is concatenated to the input (6 channels)."""
    image_channels = 6

    def forward(self, input, alpha, instance_noise=0, cond=None, **kwargs):
        x = self._trunk(torch.cat([input, cond], dim=1), alpha)
        return x.reshape(x.shape[0], 1)
