"""FiLM-SIREN decoders of the feature-volume + global-feature ("FG") family, host side.

Mirror of the reference classes ``TALLSIREN_FG`` (generators/siren.py:491-580), ``SHORTSIREN_FG``
(:583-668), ``DOUBLESIREN_FG`` (:744-827) and ``SingleSIREN_dg`` (:983-1065): same constructor
keywords, same ``forward(points, z, img_size, num_steps) -> rgb_sigma[B,N,4]``, same parameter
names (``network.{i}.layer.{weight,bias}``, ``final_layer.*``, ``mapping_network.*``) so reference
checkpoints ``load_state_dict`` strictly, same initial distributions.  The arithmetic is not
PyTorch: the trilinear lookup is ``cng_gather_points`` and the whole MLP is ``cng_film_siren_fwd``
(one fused kernel; tcgen05 bf16 or exact fp32), see include/cng_b200.h.

The config spellings ``TALLSIREN_dg`` / ``SHORTSIREN_dg`` / ``DoubleSIREN_dg``
(configs/thousand/direct_volume/dg.py:8,51,55) resolve to the same classes.
"""
from __future__ import annotations

import math
import os
from typing import List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

__all__ = ["CustomMappingNetwork", "SHORTSIREN", "FiLMLayer", "SirenLayer", "ResSirenBlock", "TALLSIREN_dRes", "TALLSIREN_dResLong", "SHORTSIREN_FRes", "TALLSIREN_FG", "SHORTSIREN_FG", "DOUBLESIREN_FG", "SingleSIREN_dg", "SHORTSIREN_F",
           "TALLSIREN_dg", "SHORTSIREN_dg", "DoubleSIREN_dg", "DOUBLESIREN_dg", "default_precision"]


def default_precision(class_default: str = "bf16") -> str:
    """Arithmetic of the fused MLP: 'bf16' / 'fp16' (tcgen05 tensor-core path, 16-bit operands, fp32 accumulate)
    or 'fp32' (exact FFMA path).  CNG_PRECISION overrides the class default; asking for bf16 on a class that is only
    offered with fp16 operands (see ``_FiLMSirenFG.precision``) selects fp16."""
    p = os.environ.get("CNG_PRECISION", class_default)
    return "fp16" if (p == "bf16" and class_default == "fp16") else p


class FiLMLayer(nn.Module):
    """Parameter holder for one ``sin(freq * (W x + b) + phase)`` layer (siren.py:146-160).

    Only ``layer`` (an ``nn.Linear``) carries state; the arithmetic runs inside the fused kernel.
    """

    def __init__(self, input_dim: int, hidden_dim: int, drop_out_prob: float = 0):
        super().__init__()
        self.layer = nn.Linear(input_dim, hidden_dim)
        self.dropout_layer = nn.Dropout(drop_out_prob)        # no parameters; kept for the module tree of siren.py:146-151
        self.drop_out_prob = drop_out_prob


SirenLayer = FiLMLayer     # siren.py:180-199: the unmodulated layer holds the same single nn.Linear


def _uniform_(linear: nn.Linear, bound: float) -> None:
    with torch.no_grad():
        linear.weight.uniform_(-bound, bound)


class _FiLMSirenFG(nn.Module):
    num_layers = 0          # FiLM layers
    freq_div = 25.0         # frequency_init(freq_div), siren.py:134-143
    sigmoid_rgb = True      # _sigmoid_rgb on the head (siren.py:579) or raw rgb (:1064)
    tensor_core_operands = "bf16"   # 16-bit operand format of the tcgen05 path that keeps this variant <= 1e-2 max-abs
    film = True             # False: plain sin(W x + b) layers, no mapping network, z is the feature volume alone
    latent = False          # True (SHORTSIREN): no feature volume, the input of layer 0 is the sample position, z is a latent vector

    def __init__(self, input_dim=3, z_dim=100, hidden_dim=256, output_dim=4, drop_out=0, device=None, **kwargs):
        super().__init__()
        self.device = device
        self.input_dim, self.z_dim, self.hidden_dim, self.output_dim = input_dim, z_dim, hidden_dim, output_dim
        self.network = nn.ModuleList(
            [FiLMLayer(input_dim if i == 0 else hidden_dim, hidden_dim, drop_out) for i in range(self.num_layers)])
        self.final_layer = nn.Linear(hidden_dim, 4)
        if self.film:
            self.mapping_network = nn.Linear(z_dim, self.num_layers * hidden_dim * 2)
        for i, film in enumerate(self.network):
            fan_in = film.layer.weight.shape[-1]
            # first_layer_film_sine_init (siren.py:40-44) overrides frequency_init on layer 0
            _uniform_(film.layer, 1.0 / fan_in if i == 0 else math.sqrt(6.0 / fan_in) / self.freq_div)
        _uniform_(self.final_layer, math.sqrt(6.0 / hidden_dim) / self.freq_div)
        self.precision = default_precision(self.tensor_core_operands)

    # -- pieces shared with ImplicitGenerator3d ------------------------------------------------
    def film_parameters(self, global_feature, batch: int = 1, device=None) -> Tuple[torch.Tensor, torch.Tensor]:
        """siren.py:550-553: freq = first half * 15 + 30, phase = second half.  [B, L*HID] each.
        Unmodulated variants (``film = False``): freq = 1, phase = 0 for ``batch`` items."""
        if not self.film:
            n = self.num_layers * self.hidden_dim
            dev = device if device is not None else self.final_layer.weight.device
            return torch.ones((batch, n), device=dev), torch.zeros((batch, n), device=dev)
        needs_grad = torch.is_grad_enabled() and (global_feature.requires_grad or self.mapping_network.weight.requires_grad)
        if global_feature.is_cuda and not needs_grad:
            # inference: the library's own kernel (batch-size independent rounding, one launch)
            return ops.film_parameters(global_feature.detach(), self.mapping_network.weight.detach(), self.mapping_network.bias.detach())
        # training: a differentiable torch op; fp32 even under the trainer's autocast (freq ~ 30 multiplies the pre-activations)
        with torch.autocast(device_type=global_feature.device.type, enabled=False):
            fo = F.linear(global_feature.float(), self.mapping_network.weight.float(), self.mapping_network.bias.float())
            half = fo.shape[-1] // 2
            return (fo[..., :half] * 15 + 30).contiguous(), fo[..., half:].contiguous()

    res_save_mask = 0       # residual blocks, see cng_film_siren_fwd_res (include/cng_b200.h)
    res_add_mask = 0

    @property
    def precision(self) -> str:
        """'bf16' | 'fp16' (tcgen05 path, 16-bit operands, fp32 accumulate) | 'fp32' (exact FFMA path).  The classes built
        with ``frequency_init(12)`` (SHORTSIREN_FG / _F / _FRes) are offered with fp16 operands only: their pre-activations
        are twice as large and bf16 operands gave 1.7e-2 max-abs, over the 1e-2 contract (BASELINE north_star); fp16 runs at
        the same tensor-core rate and is the reference's own autocast dtype."""
        return self._precision

    @precision.setter
    def precision(self, value: str) -> None:
        if value not in ops.PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(ops.PRECISIONS)}, got {value!r}")
        if value == "bf16" and self.tensor_core_operands == "fp16":
            raise ValueError(f"{type(self).__name__} is not offered with bf16 operands (1.7e-2 max-abs, over the 1e-2 contract); "
                             "use precision='fp16' (same tensor-core rate) or 'fp32'")
        self._precision = value

    def linear_layers(self) -> List[nn.Linear]:
        """The network's linear layers in execution order."""
        return [f.layer for f in self.network]

    def layer_parameters(self) -> Tuple[List[torch.Tensor], List[torch.Tensor]]:
        lin = self.linear_layers()
        return [m.weight for m in lin], [m.bias for m in lin]

    def check_dropout(self) -> None:
        """siren.py:157-159 applies ``nn.Dropout`` to a layer's output only in training mode: in eval mode (every inference
        caller) a network built with drop_out > 0 is exactly the network the fused kernels evaluate.  Training WITH dropout
        would need the mask inside the kernels (forward, recompute and backward) and is not built; no shipped config uses it
        (``dropout_ratio: 0``, configs/thousand/special.py:39)."""
        if self.training and any(getattr(m, "drop_out_prob", 0) > 0 for m in self.network):
            raise NotImplementedError("FiLM dropout > 0 in training mode is not built (call .eval(), or construct with drop_out=0)")

    def mlp(self, feat: torch.Tensor, freq: torch.Tensor, phase: torch.Tensor) -> torch.Tensor:
        """feat [B,N,C] -> rgb_sigma [B,N,4] through the fused FiLM-SIREN kernel."""
        self.check_dropout()
        ws, bs = self.layer_parameters()
        return ops.film_siren_fwd(feat, ws, bs, freq, phase, self.final_layer.weight, self.final_layer.bias,
                                  self.sigmoid_rgb, self.precision, self.res_save_mask, self.res_add_mask)

    def split_z(self, z):
        if not self.film:
            if isinstance(z, (tuple, list)):
                raise ValueError(f"{type(self).__name__} takes the feature volume alone as z (generators/siren.py:867)")
            return z, None
        if not isinstance(z, (tuple, list)) or len(z) != 2:
            raise ValueError("the FG SIREN family needs z = (feature_volume [B,C,D,H,W], global_feature [B,z_dim]) "
                             "(unet.return_global=True, generators/unet3d.py:635-638)")
        return z[0], z[1]

    def forward(self, points: torch.Tensor, z, img_size: int, num_steps: int) -> torch.Tensor:
        """points [B, N, 3] world space (N == img_size**2 * num_steps in the reference's callers;
        any N works here), z = (feature_volume, global_feature).  Returns rgb_sigma [B, N, 4]."""
        volume, global_feature = self.split_z(z)
        if torch.is_grad_enabled() and (volume.requires_grad or (global_feature is not None and global_feature.requires_grad)
                                        or any(p.requires_grad for p in self.parameters())):
            from .autograd import siren_forward_with_grad
            return siren_forward_with_grad(self, points, volume, global_feature)
        freq, phase = self.film_parameters(global_feature, volume.shape[0], volume.device)
        feat = ops.gather_points(ops.volume_to_channels_last(volume), points)
        return self.mlp(feat, freq, phase)


class CustomMappingNetwork(nn.Module):
    """generators/siren.py:55-78: z -> (frequencies, phase shifts) through three hidden Linear + LeakyReLU(0.2) layers; same module
    tree (``network.{0,2,4,6}``), kaiming_leaky_init (``:47-52``), last weight scaled by 0.25.  A [B, z_dim] x [z_dim, 256] chain per
    forward: plain torch ops (differentiable), not on the per-point path."""

    def __init__(self, z_dim: int, map_hidden_dim: int, map_output_dim: int):
        super().__init__()
        self.network = nn.Sequential(nn.Linear(z_dim, map_hidden_dim), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(map_hidden_dim, map_hidden_dim), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(map_hidden_dim, map_hidden_dim), nn.LeakyReLU(0.2, inplace=True),
                                     nn.Linear(map_hidden_dim, map_output_dim))
        for m in self.network:
            if isinstance(m, nn.Linear):
                torch.nn.init.kaiming_normal_(m.weight, a=0.2, mode="fan_in", nonlinearity="leaky_relu")
        with torch.no_grad():
            self.network[-1].weight *= 0.25

    def forward(self, z):
        fo = self.network(z)
        half = fo.shape[-1] // 2
        return fo[..., :half], fo[..., half:]


class SHORTSIREN(_FiLMSirenFG):
    """generators/siren.py:1172-1224, the generator of the default config (configs/thousand/special.py:45-51): four FiLM layers
    on the WORLD POSITION of a sample (``input_dim`` = 3, no feature volume), FiLM parameters from a latent vector ``z`` [B, z_dim]
    (PointNet encoder) through ``CustomMappingNetwork``.  On the fused kernels the positions travel as the first ``input_dim`` of
    the 32 operand channels (the rest zero) and layer 0's weight is zero-padded to [256, 32] -- layer 0 then runs in the split
    hi/lo format, i.e. the positions enter at ~fp32 precision."""
    num_layers, freq_div, sigmoid_rgb = 4, 25.0, True
    latent = True

    def __init__(self, input_dim=2, z_dim=100, hidden_dim=256, output_dim=1, drop_out=0, mapping_network="CustomMappingNetwork", device=None, **kwargs):
        nn.Module.__init__(self)
        if mapping_network != "CustomMappingNetwork":
            raise NotImplementedError(f"mapping network {mapping_network!r} (only CustomMappingNetwork is built)")
        if not 1 <= input_dim <= 32:
            raise ValueError("SHORTSIREN: input_dim must be in [1, 32]")
        self.device = device
        self.input_dim, self.z_dim, self.hidden_dim, self.output_dim = input_dim, z_dim, hidden_dim, output_dim
        self.network = nn.ModuleList([FiLMLayer(input_dim if i == 0 else hidden_dim, hidden_dim, drop_out) for i in range(self.num_layers)])
        self.final_layer = nn.Linear(hidden_dim, 4)
        self.mapping_network = CustomMappingNetwork(z_dim, 256, self.num_layers * hidden_dim * 2)
        for i, film in enumerate(self.network):
            fan_in = film.layer.weight.shape[-1]
            _uniform_(film.layer, 1.0 / fan_in if i == 0 else math.sqrt(6.0 / fan_in) / self.freq_div)
        _uniform_(self.final_layer, math.sqrt(6.0 / hidden_dim) / self.freq_div)
        self.precision = default_precision(self.tensor_core_operands)

    def split_z(self, z):
        if isinstance(z, (tuple, list)) or z.dim() != 2:
            raise ValueError("SHORTSIREN takes a latent vector z [B, z_dim] (generators/siren.py:1206-1208)")
        return None, z

    def film_parameters(self, z, batch: int = 1, device=None):
        with torch.autocast(device_type=z.device.type, enabled=False):
            freq, phase = self.mapping_network(z.float())
            return (freq * 15 + 30).contiguous(), phase.contiguous()

    def layer_parameters(self):
        lin = self.linear_layers()
        ws = [F.pad(lin[0].weight, (0, 32 - self.input_dim))] + [m.weight for m in lin[1:]]      # positions occupy operand channels 0..input_dim-1
        return ws, [m.bias for m in lin]

    def point_features(self, points: torch.Tensor) -> torch.Tensor:
        """[B, N, input_dim] positions -> the kernels' 32-channel operand rows [B, N, 32]."""
        return F.pad(points.float(), (0, 32 - points.shape[-1])).contiguous()

    def forward(self, input: torch.Tensor, z, *args) -> torch.Tensor:
        _, latent = self.split_z(z)
        freq, phase = self.film_parameters(latent, input.shape[0], input.device)
        feat = self.point_features(input)
        if torch.is_grad_enabled() and (latent.requires_grad or any(p.requires_grad for p in self.parameters())):
            from .autograd import _mlp
            return _mlp(self, feat, freq, phase)
        return self.mlp(feat, freq, phase)


class TALLSIREN_FG(_FiLMSirenFG):
    num_layers, freq_div, sigmoid_rgb = 8, 25.0, True


class SHORTSIREN_FG(_FiLMSirenFG):
    num_layers, freq_div, sigmoid_rgb = 4, 12.0, True
    # frequency_init(12) doubles the pre-activations: bf16 operands give 1.4e-2 max-abs on random-init weights, fp16
    # operands (same tensor-core rate, the reference's own autocast dtype) 2e-3
    tensor_core_operands = "fp16"


class DOUBLESIREN_FG(_FiLMSirenFG):
    num_layers, freq_div, sigmoid_rgb = 2, 12.0, True


class SingleSIREN_dg(_FiLMSirenFG):
    num_layers, freq_div, sigmoid_rgb = 1, 25.0, False


class SHORTSIREN_F(_FiLMSirenFG):
    """generators/siren.py:830-904: feature volume only, four plain ``sin(W x + b)`` layers (SirenLayer), no FiLM, no
    mapping network; ``z`` is the feature volume.  Runs on the same kernels with freq = 1, phase = 0."""
    num_layers, freq_div, sigmoid_rgb, film = 4, 12.0, True, False
    tensor_core_operands = "fp16"


class ResSirenBlock(nn.Module):
    """Parameter holder of ``y = sin(x + fc2(sin(fc1 x)))`` (siren.py:218-230); the arithmetic runs inside the fused kernel."""

    def __init__(self, hidden_dim: int):
        super().__init__()
        self.fc1 = nn.Linear(hidden_dim, hidden_dim)
        self.fc2 = nn.Linear(hidden_dim, hidden_dim)


class _ResSiren(_FiLMSirenFG):
    """Feature-volume-only decoders built from ``SirenLayer`` and ``ResSirenBlock`` (siren.py:333-488, 906-979): SirenLayer,
    ``num_blocks`` residual blocks, SirenLayer, ``nn.Linear(hidden, 4)`` head.  2 + 2 * num_blocks linear layers on the fused
    kernels (freq = 1, phase = 0); a block's input is kept in fp32 next to the kernel's 16-bit operand tile and added to its
    second layer's pre-activation (``cng_film_siren_fwd_res``).  State-dict keys as in the reference: ``network.0.layer.*``,
    ``network.{1..num_blocks}.fc{1,2}.*``, ``network.{num_blocks+1}.layer.*``, ``final_layer.*``.  The backward recomputes
    through the training-mode kernel like the FG family; the kept activation receives the adding layer's dz on top of its own
    gradient (``generators/autograd.py``)."""
    film = False
    num_blocks = 0
    input_from_z_dim = False          # TALLSIREN_dRes / _dResLong read z_dim features (``input_dim = z_dim``, siren.py:355, :433)

    def __init__(self, input_dim=3, z_dim=100, hidden_dim=256, output_dim=4, drop_out=0, device=None, **kwargs):
        nn.Module.__init__(self)
        if self.input_from_z_dim:
            input_dim = z_dim
        self.device = device
        self.input_dim, self.z_dim, self.hidden_dim, self.output_dim = input_dim, z_dim, hidden_dim, output_dim
        self.network = nn.ModuleList([SirenLayer(input_dim, hidden_dim, drop_out)] + [ResSirenBlock(hidden_dim) for _ in range(self.num_blocks)]
                                     + [SirenLayer(hidden_dim, hidden_dim, drop_out)])
        self.final_layer = nn.Linear(hidden_dim, 4)
        for i, lin in enumerate(self.linear_layers()):
            fan_in = lin.weight.shape[-1]
            # network.apply(frequency_init(f)) then network[0].apply(first_layer_film_sine_init), siren.py:372-375
            _uniform_(lin, 1.0 / fan_in if i == 0 else math.sqrt(6.0 / fan_in) / self.freq_div)
        _uniform_(self.final_layer, math.sqrt(6.0 / hidden_dim) / self.freq_div)
        self.precision = default_precision(self.tensor_core_operands)

    def linear_layers(self) -> List[nn.Linear]:
        n = self.network
        out = [n[0].layer]
        for b in range(1, 1 + self.num_blocks):
            out += [n[b].fc1, n[b].fc2]
        return out + [n[1 + self.num_blocks].layer]


def _res_masks(num_blocks: int) -> Tuple[int, int]:
    """(save, add): layer 0 and every fc2 keep their output; every fc2 (layers 2, 4, ...) adds the kept one."""
    save = 1 | sum(1 << (2 * b) for b in range(1, num_blocks))
    add = sum(1 << (2 * b) for b in range(1, num_blocks + 1))
    return save, add


class TALLSIREN_dRes(_ResSiren):
    """siren.py:333-408 (configs/thousand/direct_volume/dRes.py): two residual blocks, raw head."""
    num_blocks, num_layers, freq_div, sigmoid_rgb, input_from_z_dim = 2, 6, 25.0, False, True
    res_save_mask, res_add_mask = _res_masks(2)


class TALLSIREN_dResLong(_ResSiren):
    """siren.py:411-488: four residual blocks, raw head."""
    num_blocks, num_layers, freq_div, sigmoid_rgb, input_from_z_dim = 4, 10, 25.0, False, True
    res_save_mask, res_add_mask = _res_masks(4)


class SHORTSIREN_FRes(_ResSiren):
    """siren.py:906-979: one residual block, frequency_init(12), sigmoid on rgb, ``input_dim`` features."""
    num_blocks, num_layers, freq_div, sigmoid_rgb = 1, 4, 12.0, True
    res_save_mask, res_add_mask = _res_masks(1)
    tensor_core_operands = "fp16"


# config spellings
TALLSIREN_dg = TALLSIREN_FG
SHORTSIREN_dg = SHORTSIREN_FG
DoubleSIREN_dg = DOUBLESIREN_FG
DOUBLESIREN_dg = DOUBLESIREN_FG


# the decoders whose MLP is a library (PyTorch) op chain rather than the fused kernels: TALLSIREN (per-point FiLM), TALLSIREN_dgx,
# SHORTSIREN_FG_Pyrmd -- imported here so that ``getattr(siren, siren_type)`` (generators.py:15) resolves every reference class
from .siren_library import PointFeaturesMappingNetwork, PointwiseFiLMLayer, SHORTSIREN_FG_Pyrmd, TALLSIREN, TALLSIREN_dgx  # noqa: E402,F401
