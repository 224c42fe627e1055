"""Latent I2SB super-resolution sampler with the call surface of
/root/reference/afldm/pipelines/i2sb_pipeline.py:15-78 (driven by scripts/shift_ldm_sr.py): encode the degraded
image with the alias-free VAE, run the bridge from t_max down (the loop stops one step early, :48-49), decode."""
from __future__ import annotations

from typing import Optional

import torch

from .. import ops
from ..schedulers.i2sb import I2SBScheduler
from .ldm_pipeline import ImagePipelineOutput, MyLDMPipeline


class I2SBLDMPipeline(MyLDMPipeline):
    SCHEDULER_CLS = I2SBScheduler

    def __init__(self, vae, unet, scheduler: I2SBScheduler):
        super().__init__(vae, unet, scheduler)

    @torch.no_grad()
    def bridge(self, latents: torch.Tensor, num_inference_steps: int = 50, is_ode: bool = False, generator=None):
        """The sampling loop of i2sb_pipeline.py:45-56 on latents [B,C,H,W]."""
        self.scheduler.set_timesteps(num_inference_steps)
        latents = latents.to(device=self.device, dtype=torch.float32)
        for i, t in enumerate(self.progress_bar(self.scheduler.timesteps)):
            if i == num_inference_steps - 1:
                break
            eps = self.unet(self.scheduler.scale_model_input(latents, t), int(t)).sample
            latents = self.scheduler.step(eps, int(t), latents, is_ode=is_ode, generator=generator).prev_sample
        return ops.to_nchw_contiguous(ops.nhwc(latents))

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, generator=None, is_ode: bool = False, num_inference_steps: int = 50,
                 output_type: Optional[str] = "pil", return_dict: bool = True, **kwargs):
        """``images``: [B,3,H,W] in [-1, 1] (the reference's VaeImageProcessor.preprocess maps [0,1] inputs there)."""
        x = images.to(device=self.device, dtype=torch.float32)
        latents = self.vae.encode(x).latent_dist.sample(generator) * self.vae.config.scaling_factor
        latents = self.bridge(latents, num_inference_steps, is_ode, generator)
        if output_type == "latent":
            return latents
        image = self.vae.decode(latents / self.vae.config.scaling_factor).sample
        if output_type != "pt":
            image = (image / 2 + 0.5).clamp(0, 1).cpu().permute(0, 2, 3, 1).numpy()
            if output_type == "pil":
                from PIL import Image
                image = [Image.fromarray((im * 255).round().astype("uint8")) for im in image]
            return ImagePipelineOutput(images=image) if return_dict else (image,)
        return image
