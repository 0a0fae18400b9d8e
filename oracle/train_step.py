"""CPU restatement of the reference's GAN train step (BASELINE config 3) -- TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

``Trainer`` itself cannot be imported here (``utils.py`` needs ``configs.thesis``, OpenEXR, lpips ...),
so this harness follows ``Trainer.train_discriminator`` (``utils.py:743-842``) and ``Trainer.train_generator``
(``utils.py:621-741``) statement by statement around whatever three modules it is given:

  * ``tests/golden/make_golden.py`` runs it with the REFERENCE's ``ImplicitGenerator3d``, ``unet3d.UNet3D`` and
    ``ProgressiveDiscriminator`` (imported from /root/reference) and records losses and gradient norms
    (``tests/golden/train_step.npz``) -- that pins this harness and the fixtures to the reference's own modules;
  * the CPU tests run it with this repository's U-Net and discriminator (and the oracle's renderer as the generator);
  * the GPU tests compare the product's ``training.GanTrainStep`` against the same fixture.

CPU / fp32: autocast and the GradScaler are identities (``GradScaler(enabled=False)`` returns its argument from
``scale`` and 1.0 from ``get_scale``), so the statements that use them are kept but inert.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import nerf_path


def fill_params(module: nn.Module, seed: int) -> None:
    """Deterministic, init-order-independent parameters: parameter i (state-dict order) is drawn from its own generator.
    Weights ~ N(0, 1/fan_in), biases ~ N(0, 0.1^2), normalisation scales 1 + N(0, 0.1^2)."""
    with torch.no_grad():
        for i, (name, p) in enumerate(module.named_parameters()):
            g = torch.Generator().manual_seed(seed * 100003 + i)
            r = torch.randn(p.shape, generator=g)
            if p.dim() > 1:
                fan_in = p[0].numel()
                p.copy_(r / fan_in ** 0.5)
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1 + 0.1 * r)
            else:
                p.copy_(0.1 * r)


class OracleGenerator(nn.Module):
    """``ImplicitGenerator3d`` stand-in for CPU runs of the harness: parameters under the reference's state-dict names,
    forward = the oracle's differentiable render with replayed draws (``metadata["draws"]``)."""

    def __init__(self, siren_type: str, state: Dict[str, torch.Tensor]):
        super().__init__()
        self.siren_type = siren_type
        self.names = list(state)
        self.params = nn.ParameterList([nn.Parameter(v.clone()) for v in state.values()])
        self.step = 0

    def state(self) -> Dict[str, torch.Tensor]:
        return dict(zip(self.names, self.params))

    def forward(self, z, cam2worlds, **md):
        meta = {k: v for k, v in md.items() if k != "draws"}
        out = nerf_path.render_with_grad(self.state(), self.siren_type, z, cam2worlds, md["draws"], **meta)
        return out["pixels"], out["depth"]


def tiny_config() -> Dict:
    """The curriculum of the fixture: ``configs/thousand/default.py`` + ``special.py`` values at toy sizes."""
    return {
        "img_size": 16, "num_steps": 4, "fov": 49.134342641202636, "ray_start": 0.25, "ray_end": 1.95, "hierarchical_sample": True,
        "clamp_mode": "relu", "white_back": True, "nerf_noise": 0.5, "batch_split": 1, "r1_lambda": 10, "grad_clip": 1,
        "betas": (0.0, 0.9), "weight_decay": 0, "gen_lr": 1e-5, "disc_lr": 1e-4, "enc_lr": 2e-5, "photo_loss": True, "depth_loss": False,
        "depth_loss_weight": 1, "enable_discriminator": True, "random_gen_img": False, "cam_r_start": 0.7, "cam_r_end": 1.5,
        "fade_steps": 2000,
    }


def tiny_curriculum() -> Dict:
    """Four stages as in configs/thousand/default.py:11-60 (32 -> 64 -> 128 -> 128 pixels) plus global keys."""
    return {0: {"batch_size": 32, "img_size": 32, "num_steps": 48}, 5000: {"batch_size": 24, "img_size": 64, "num_steps": 48},
            15000: {"batch_size": 4, "img_size": 128, "num_steps": 48}, 25000: {"batch_size": 4, "img_size": 128, "num_steps": 64},
            "fade_steps": 2000, "fov": 30}


TINY_UNET = dict(in_channels=4, out_channels=32, f_maps=8, num_levels=2, is_segmentation=False, final_sigmoid=False, return_global=True)
TINY_SIREN, TINY_ZDIM, TINY_BATCH, TINY_VOXEL = "DOUBLESIREN_FG", 16, 2, 8


def tiny_sample(seed: int = 0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    B, V, img = TINY_BATCH, TINY_VOXEL, tiny_config()["img_size"]
    occ = (torch.rand((B, 1, V, V, V), generator=g) < 0.3).float()
    voxel = torch.cat([occ, torch.rand((B, 3, V, V, V), generator=g) * occ], dim=1)
    import numpy as np
    origins = nerf_path.random_camera_origins(B, 0.7, 1.5, "y", np.random.RandomState(seed))
    return {"img": torch.rand((B, 3, img, img), generator=g) * 2 - 1, "voxel": voxel, "cam2world": nerf_path.look_at_cam2world(origins)}


def tiny_draws(seed: int = 1):
    md = tiny_config()
    return nerf_path.draw_randoms(TINY_BATCH, md["img_size"], md["num_steps"], True, torch.Generator().manual_seed(seed))


class RefTrainStep:
    """One optimisation step of the reference trainer on the given modules; records what the tests compare."""

    def __init__(self, generator, encoder, discriminator, metadata: Dict, alpha: float = 1.0):
        self.generator_ddp, self.encoder_ddp, self.discriminator_ddp = generator, encoder, discriminator
        self.metadata = dict(metadata)
        md = self.metadata
        self.device = torch.device("cpu")
        self.alpha = alpha
        adam = lambda m, lr: torch.optim.Adam(m.parameters(), lr=lr, betas=md["betas"], weight_decay=md["weight_decay"])   # utils.py:327-338, 353-358, 392-397
        self.optimizer_G, self.optimizer_D, self.optimizer_E = adam(generator, md["gen_lr"]), adam(discriminator, md["disc_lr"]), adam(encoder, md["enc_lr"])
        self.scaler = torch.amp.GradScaler("cpu", enabled=False)
        self.record: Dict[str, float] = {}

    def train_discriminator(self, sample):
        md = self.metadata
        imgs = sample["img"]                                                                     # utils.py:747
        split_batch_size = imgs.shape[0] // md["batch_split"]                                    # :753
        real_imgs = imgs.to(self.device)                                                         # :756
        voxels = sample["voxel"].to(self.device)                                                 # :758
        with torch.no_grad():                                                                    # :761
            cam2worlds = sample["cam2world"].to(self.device)                                     # :772 (random_gen_img False)
            gen_imgs: List[torch.Tensor] = []
            for split in range(md["batch_split"]):                                               # :775
                subset_z = self.encoder_ddp(voxels[split * split_batch_size:(split + 1) * split_batch_size])   # :777
                subset_cam = cam2worlds[split * split_batch_size:(split + 1) * split_batch_size]                # :787
                gen_img, gen_depth = self.generator_ddp(subset_z, subset_cam, **md)             # :790
                gen_imgs.append(gen_img)
            gen_imgs = torch.cat(gen_imgs, dim=0)                                                # :799
        real_imgs = real_imgs.clone()
        real_imgs.requires_grad = True                                                           # :802
        r_preds = self.discriminator_ddp(real_imgs, self.alpha, cond=[None] * imgs.shape[0], **md)     # :803
        if md["r1_lambda"] > 0:                                                                  # :807
            grad_real = torch.autograd.grad(outputs=self.scaler.scale(r_preds.sum()), inputs=real_imgs, create_graph=True)
            inv_scale = 1.0 / self.scaler.get_scale()
            grad_real = [p * inv_scale for p in grad_real][0]
            grad_penalty = (grad_real.view(grad_real.size(0), -1).norm(2, dim=1) ** 2).mean()    # :818
            grad_penalty = 0.5 * md["r1_lambda"] * grad_penalty
        else:
            grad_penalty = 0
        g_preds = self.discriminator_ddp(gen_imgs, self.alpha, cond=[None] * imgs.shape[0], **md)      # :825
        d_loss = F.softplus(g_preds).mean() + F.softplus(-r_preds).mean() + grad_penalty         # :829
        self.record["d_loss"] = d_loss.item()
        self.record["grad_penalty"] = float(grad_penalty)
        self.optimizer_D.zero_grad()                                                             # :836
        self.scaler.scale(d_loss).backward()
        self.scaler.unscale_(self.optimizer_D)
        self.record["norm_D"] = float(torch.nn.utils.clip_grad_norm_(self.discriminator_ddp.parameters(), md["grad_clip"]))   # :839
        self.scaler.step(self.optimizer_D)                                                       # :842

    def train_generator(self, sample):
        md = self.metadata
        imgs = sample["img"].to(self.device)                                                     # :625
        cam2worlds = sample["cam2world"].to(self.device)
        voxels = sample["voxel"].to(self.device)
        split_batch_size = imgs.shape[0] // md["batch_split"]                                    # :639
        g_buf = p_buf = 0.0
        for split in range(md["batch_split"]):                                                   # :643
            sl = slice(split * split_batch_size, (split + 1) * split_batch_size)
            subset_z = self.encoder_ddp(voxels[sl])                                              # :646
            gen_imgs, gen_depths = self.generator_ddp(subset_z, cam2worlds[sl], **md)            # :659
            if md["enable_discriminator"]:
                g_preds = self.discriminator_ddp(gen_imgs, self.alpha, cond=[None] * split_batch_size, **md)   # :665
                loss_G = F.softplus(-g_preds).mean()                                             # :672
            else:
                loss_G = torch.zeros(1)
            photometry_loss = ((imgs[sl] - gen_imgs) ** 2).mean() if md["photo_loss"] else torch.zeros(1)     # :676, utils.py:102-104
            loss = loss_G + photometry_loss                                                      # :701 (depth / z_reg terms are zero here)
            g_buf += loss_G.item()
            p_buf += photometry_loss.item()
            self.scaler.scale(loss).backward()                                                   # :711
        self.record["g_loss"], self.record["photo_loss"] = g_buf / md["batch_split"], p_buf / md["batch_split"]
        self.scaler.unscale_(self.optimizer_G)                                                   # :726
        self.record["norm_G"] = float(torch.nn.utils.clip_grad_norm_(self.generator_ddp.parameters(), md.get("grad_clip", 0.3)))
        self.scaler.step(self.optimizer_G)
        self.optimizer_G.zero_grad()
        self.scaler.unscale_(self.optimizer_E)                                                   # :734
        self.record["norm_E"] = float(torch.nn.utils.clip_grad_norm_(self.encoder_ddp.parameters(), md.get("grad_clip", 0.3)))
        self.scaler.step(self.optimizer_E)
        self.optimizer_E.zero_grad()
        self.scaler.update()                                                                     # :741

    def step(self, sample) -> Dict[str, float]:
        """train.py:100-105: discriminator first, then generator + encoder."""
        self.record = {}
        self.train_discriminator(sample)
        self.train_generator(sample)
        return dict(self.record)
