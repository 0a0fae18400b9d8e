"""The GAN train step around the rendering path (BASELINE config 3).

Follows ``Trainer.train_discriminator`` (``utils.py:743-842``) and ``Trainer.train_generator`` (``utils.py:621-741``) in
the order ``train.py:96-105`` runs them, for the voxel-conditioned setting (``dataset.load_voxel``): the 3D U-Net
encodes the voxel grid into ``(feature volume, global feature)``, the generator renders it, the progressive
discriminator scores the images.  Kept from the reference: the losses (non-saturating logistic + R1 on real images +
photometric / depth terms), ``GradScaler`` usage and its unscale / clip / step order, ``batch_split`` micro-batching, the
random cameras of the discriminator step, ``set_alpha`` (fade-in ``alpha`` and the ``nerf_noise`` schedule,
``utils.py:610-618``).  Out of scope here as in the reference trainer's other 900 lines: datasets, checkpoints, logging,
FID.

Differences that do not change the mathematics:
  * one process per GPU with NCCL; under DDP the micro-batch backward passes of one optimizer step run inside
    ``no_sync()`` except the last, so gradients cross NVLink once per step (the reference all-reduces after every
    micro-batch backward, ``utils.py:711``);
  * losses are accumulated on the device and read back once per step (the reference calls ``.item()`` three times per
    micro-batch, ``utils.py:707-709``, a host synchronisation each);
  * the encoder emits the feature volume in the gather kernel's NDHWC layout (``generators/unet3d.py``);
  * in the generator step the discriminator only relays d loss / d image: it is called as the bare module with frozen
    parameters, so its weight gradients are neither computed nor all-reduced (the reference computes them, lets DDP
    average them and zeroes them at the next discriminator step, ``utils.py:836``).
"""
from __future__ import annotations

import contextlib
import os
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from .generators.volumetric_rendering import create_cam2world_matrix, sample_camera_positions


def loss_mse(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """utils.py:102-105"""
    return ((x - y) ** 2).mean()


def loss_depth(gt: torch.Tensor, preds: torch.Tensor) -> torch.Tensor:
    """MSE over the ground truth's foreground pixels (utils.py:96-99)."""
    mask = gt != 0
    return ((gt[mask] - preds[mask]) ** 2).mean()


def _detach(z):
    """detach() through the (volume, global feature) tuples / pyramids an encoder may return"""
    if isinstance(z, (tuple, list)):
        return type(z)(_detach(t) for t in z)
    return z.detach()


class GanTrainStep:
    """Owns the three optimizers and the GradScaler; ``step(sample)`` = one discriminator update + one generator/encoder
    update on ``sample = {"img": [B,3,H,W], "voxel": [B,4,V,V,V], "cam2world": [B,4,4] (, "depth": [B,H,W])}``.

    ``metadata`` is the flat curriculum dict of the reference (``curriculums.extract_metadata``): it is passed whole to the
    generator and the discriminator, which ignore the keys they do not know."""

    def __init__(self, generator, encoder, discriminator, metadata: Dict, device, *, amp: bool = True, amp_dtype=torch.float16,
                 ddp: bool = False, local_rank: int = 0, fade_steps: Optional[int] = None, last_upsample_step: int = 0,
                 curriculum: Optional[Dict] = None):
        self.device = torch.device(device)
        self.metadata = dict(metadata)
        self.generator, self.encoder, self.discriminator = generator, encoder, discriminator
        self.amp, self.amp_dtype = amp and self.device.type == "cuda", amp_dtype
        ids = [local_rank] if self.device.type == "cuda" else None          # gloo / CPU (tests): no device ids
        wrap = (lambda m, unused: torch.nn.parallel.DistributedDataParallel(m, device_ids=ids, find_unused_parameters=unused)) \
            if ddp else (lambda m, unused: m)
        # utils.py:321-326, 344-348, 385-389: generator and discriminator with find_unused_parameters, the encoder without
        self.generator_ddp = wrap(generator, True)
        self.encoder_ddp = wrap(encoder, False)
        self.discriminator_ddp = wrap(discriminator, True) if discriminator is not None else None
        md = self.metadata
        # utils.py:327-338, 353-358, 392-397; fused=True on CUDA: one multi-tensor kernel per optimizer step instead of a
        # dozen foreach launches (the 8-GPU step at 4 images per GPU is launch-bound), same update rule
        adam = lambda params, lr: torch.optim.Adam(params, lr=lr, betas=tuple(float(b) for b in md.get("betas", (0, 0.9))),
                                                   weight_decay=md.get("weight_decay", 0), fused=self.device.type == "cuda")
        self.optimizer_G = adam(self.generator_ddp.parameters(), md["gen_lr"])
        self.optimizer_E = adam(self.encoder_ddp.parameters(), md["enc_lr"])
        self.optimizer_D = adam(self.discriminator_ddp.parameters(), md["disc_lr"]) if discriminator is not None else None
        self.scaler = torch.amp.GradScaler(self.device.type, enabled=self.amp and amp_dtype == torch.float16)
        self.fade_steps = fade_steps if fade_steps is not None else md.get("fade_steps", 2000)
        self.last_upsample_step = last_upsample_step
        self.curriculum = curriculum             # optional: the reference's curriculum dict; ``set_alpha`` then reads the stage start from it
        self.alpha = 1.0
        self.ddp = ddp
        # step(): one encoder forward serves the D step and the G/E step (see there); CNG_SHARE_ENCODER=0 restores the two passes
        self.share_encoder_forward = os.environ.get("CNG_SHARE_ENCODER", "1") != "0"
        self._shared_z = None
        self._encoder_deterministic = None
        self.losses: Dict[str, torch.Tensor] = {}
        self.grad_norms: Dict[str, torch.Tensor] = {}          # total norms returned by clip_grad_norm_ (device scalars)
        if hasattr(generator, "set_device"):
            generator.set_device(self.device)

    # ------------------------------------------------------------------------------------------------------------
    def _encoder_is_deterministic(self) -> bool:
        if self._encoder_deterministic is None:
            bad = (torch.nn.modules.batchnorm._BatchNorm, torch.nn.modules.dropout._DropoutNd)
            self._encoder_deterministic = not any(isinstance(m, bad) and (not isinstance(m, torch.nn.modules.dropout._DropoutNd) or m.p > 0)
                                                  for m in self.encoder_ddp.modules())
        return self._encoder_deterministic

    def _autocast(self):
        return torch.autocast(self.device.type, dtype=self.amp_dtype, enabled=self.amp)

    def _no_sync(self, module, last: bool):
        return module.no_sync() if (self.ddp and not last) else contextlib.nullcontext()

    def set_alpha(self) -> None:
        """utils.py:610-618"""
        step = getattr(self.generator, "step", 0)
        if self.curriculum is not None:
            from . import curriculums
            self.last_upsample_step = curriculums.last_upsample_step(self.curriculum, step)
        self.alpha = min(1, (step - self.last_upsample_step) / self.fade_steps)
        self.metadata["nerf_noise"] = max(0, 1.0 - step / 5000.0)

    # ------------------------------------------------------------------------------------------------------------
    def train_discriminator(self, sample: Dict) -> None:
        """utils.py:743-842"""
        md = self.metadata
        imgs = sample["img"]
        B = imgs.shape[0]
        splits = md.get("batch_split", 1)
        sb = B // splits
        with self._autocast():
            real_imgs = imgs.to(self.device, non_blocking=True)
            voxels = sample["voxel"].to(self.device, non_blocking=True)
            with torch.no_grad():
                if md.get("random_gen_img", True):
                    origins = sample_camera_positions(self.device, up_direction="y", cam_r_start=md["cam_r_start"], cam_r_end=md["cam_r_end"], n=B)
                    cam2worlds = create_cam2world_matrix(origins, "y", self.device)
                else:
                    cam2worlds = sample["cam2world"].to(self.device)
                gen_imgs = []
                for s in range(splits):
                    z = _detach(self._shared_z) if self._shared_z is not None else self.encoder_ddp(voxels[s * sb:(s + 1) * sb])
                    gen_img, _ = self.generator_ddp(z, cam2worlds[s * sb:(s + 1) * sb], **md)
                    gen_imgs.append(gen_img)
                gen_imgs = torch.cat(gen_imgs, dim=0)
            real_imgs = real_imgs.detach().requires_grad_(True)
            r_preds = self.discriminator_ddp(real_imgs, self.alpha, cond=None, **md)
        r1 = md.get("r1_lambda", 0)
        if r1 > 0:
            from .discriminators.discriminators import r1_pass
            with r1_pass():
                grad_real = torch.autograd.grad(outputs=self.scaler.scale(r_preds.sum()), inputs=real_imgs, create_graph=True)[0]
            if self.scaler.is_enabled():
                # un-scale on the device (GradScaler keeps its scale in a device tensor): get_scale() would be a host
                # synchronisation in the middle of the step
                grad_real = grad_real * self.scaler._scale.reciprocal().to(grad_real.dtype)
        with self._autocast():
            if r1 > 0:
                grad_penalty = 0.5 * r1 * (grad_real.reshape(grad_real.size(0), -1).norm(2, dim=1) ** 2).mean()
            else:
                grad_penalty = 0
            g_preds = self.discriminator_ddp(gen_imgs, self.alpha, cond=None, **md)
            d_loss = F.softplus(g_preds).mean() + F.softplus(-r_preds).mean() + grad_penalty
        self.losses["d_loss"] = d_loss.detach()
        self.optimizer_D.zero_grad()
        self.scaler.scale(d_loss).backward()
        self.scaler.unscale_(self.optimizer_D)
        self.grad_norms["D"] = torch.nn.utils.clip_grad_norm_(self.discriminator_ddp.parameters(), md["grad_clip"])
        self.scaler.step(self.optimizer_D)

    # ------------------------------------------------------------------------------------------------------------
    def train_generator(self, sample: Dict) -> None:
        """utils.py:621-741"""
        md = self.metadata
        imgs = sample["img"].to(self.device, non_blocking=True)
        cam2worlds = sample["cam2world"].to(self.device)
        voxels = sample["voxel"].to(self.device, non_blocking=True)
        depths = sample.get("depth")
        splits = md.get("batch_split", 1)
        sb = imgs.shape[0] // splits
        use_d = md.get("enable_discriminator", True) and self.discriminator_ddp is not None
        zero = torch.zeros((), device=self.device)
        acc = {"g_loss": zero.clone(), "photo_loss": zero.clone(), "depth_loss": zero.clone()}
        # The generator loss needs d loss / d image through the discriminator, not the discriminator's weight gradients
        # (the reference computes, all-reduces and then discards them): score with the bare module and frozen parameters.
        d_params = [p for p in self.discriminator.parameters() if p.requires_grad] if use_d else []
        for p in d_params:
            p.requires_grad_(False)
        try:
            self._generator_micro_batches(md, imgs, cam2worlds, voxels, depths, splits, sb, use_d, zero, acc)
        finally:
            for p in d_params:
                p.requires_grad_(True)
        for k, v in acc.items():
            self.losses[k] = v / splits
        clip = md.get("grad_clip", 0.3)
        self.scaler.unscale_(self.optimizer_G)
        self.grad_norms["G"] = torch.nn.utils.clip_grad_norm_(self.generator_ddp.parameters(), clip)
        self.scaler.step(self.optimizer_G)
        self.optimizer_G.zero_grad()
        self.scaler.unscale_(self.optimizer_E)
        self.grad_norms["E"] = torch.nn.utils.clip_grad_norm_(self.encoder_ddp.parameters(), clip)
        self.scaler.step(self.optimizer_E)
        self.optimizer_E.zero_grad()
        self.scaler.update()

    def _generator_micro_batches(self, md, imgs, cam2worlds, voxels, depths, splits, sb, use_d, zero, acc) -> None:
        for s in range(splits):
            last = s == splits - 1
            sl = slice(s * sb, (s + 1) * sb)
            with self._no_sync(self.generator_ddp, last), self._no_sync(self.encoder_ddp, last):
                with self._autocast():
                    z = self._shared_z if self._shared_z is not None else self.encoder_ddp(voxels[sl])
                    gen_imgs, gen_depths = self.generator_ddp(z, cam2worlds[sl], **md)
                    if use_d:
                        g_preds = self.discriminator(gen_imgs, self.alpha, cond=None, **md)
                        loss_G = F.softplus(-g_preds).mean()
                    else:
                        loss_G = zero
                    photo = loss_mse(imgs[sl], gen_imgs) if md.get("photo_loss", False) else zero
                    depth = loss_depth(depths[sl].to(self.device), gen_depths) if md.get("depth_loss", False) else zero
                    loss = loss_G + photo + depth * md.get("depth_loss_weight", 1)
                acc["g_loss"] += loss_G.detach().float()
                acc["photo_loss"] += photo.detach().float()
                acc["depth_loss"] += depth.detach().float()
                self.scaler.scale(loss).backward()

    # ------------------------------------------------------------------------------------------------------------
    def step(self, sample: Dict) -> Dict[str, torch.Tensor]:
        """train.py:92-105 and :122-125: modes, ``set_alpha``, D step, G step, step counters.  Returns the step's losses as
        device scalars (read them with ``.item()`` when needed)."""
        self._steps_done = getattr(self, "_steps_done", 0) + 1
        if self.scaler.is_enabled() and self._steps_done % 64 == 1 and self.scaler.get_scale() < 1:      # (a host sync: not every step)
            self.scaler.update(1.0)
        self.generator_ddp.train()
        self.encoder_ddp.train()
        use_d = self.metadata.get("enable_discriminator", True) and self.discriminator_ddp is not None
        if use_d:
            self.discriminator_ddp.train()
        self.set_alpha()
        # The reference encodes the voxels twice per step -- without grad for the D step's fake images (utils.py:771-775), with
        # grad for the G/E step (:652-655).  The encoder's weights do not change in between (only optimizer_D steps) and a
        # deterministic encoder (no dropout, no batch statistics) returns the same bits both times: encode once, with grad, and
        # give the D step the detached result.  Micro-batched steps (batch_split > 1) keep the two passes.
        self._shared_z = None
        if use_d and self.share_encoder_forward and self.metadata.get("batch_split", 1) == 1 and self._encoder_is_deterministic():
            with self._autocast():
                self._shared_z = self.encoder_ddp(sample["voxel"].to(self.device, non_blocking=True))
        try:
            if use_d:
                self.train_discriminator(sample)
            self.train_generator(sample)
        finally:
            self._shared_z = None
        self.generator.step = getattr(self.generator, "step", 0) + 1
        if use_d:
            self.discriminator.step = getattr(self.discriminator, "step", 0) + 1
        return dict(self.losses)
