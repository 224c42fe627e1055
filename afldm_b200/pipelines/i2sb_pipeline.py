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
    def bridge(self, latents: torch.Tensor, num_inference_steps: int = 50, is_ode: bool = False, generator=None,
               use_cuda_graph: bool = True):
        """The sampling loop of i2sb_pipeline.py:45-56 on latents [B,C,H,W].  The update is x <- x + c_eps(t) eps
        (+ sqrt(var) noise when not ``is_ode``), so the UNet evaluation and the update replay the same captured step as
        the DDIM loop (coefficients [1, c_eps] read from a device tensor); the noise term is added between replays."""
        from .ldm_pipeline import graph_capturable
        self.scheduler.set_timesteps(num_inference_steps)
        latents = latents.to(device=self.device, dtype=torch.float32)
        ts = [int(t) for t in self.scheduler.timesteps][:num_inference_steps - 1]      # the loop stops one step early (:48-49)
        if not use_cuda_graph or not graph_capturable(self.unet):
            for t in self.progress_bar(ts):
                eps = self.unet(self.scheduler.scale_model_input(latents, t), t).sample
                latents = self.scheduler.step(eps, t, latents, is_ode=is_ode, generator=generator).prev_sample
            return ops.to_nchw_contiguous(ops.nhwc(latents))
        g = self.graphed(latents.shape[0], size=latents.shape[-1])
        coefs = [self.scheduler.coefficients(t) for t in ts]
        table = torch.tensor([[1.0, float(c)] for c, _ in coefs], dtype=torch.float32, device=self.device)
        tt = torch.tensor(ts, dtype=torch.float32, device=self.device)[:, None].expand(-1, latents.shape[0]).contiguous()
        g.x.copy_(ops.nhwc(latents))
        for i, t in enumerate(self.progress_bar(ts)):
            g.t.copy_(tt[i])
            g.coef.copy_(table[i])
            g.replay()
            if t > 0 and not is_ode:
                gdev = generator.device if generator is not None else torch.device("cpu")
                noise = torch.randn(latents.shape, generator=generator, device=gdev, dtype=torch.float32).to(self.device)
                ops.axpby(g.x, ops.nhwc(noise), 1.0, float(coefs[i][1]), out=g.x)
        return ops.to_nchw_contiguous(g.x)

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, generator=None, is_ode: bool = False, num_inference_steps: int = 50,
                 output_type: Optional[str] = "pil", return_dict: bool = True, **kwargs):
        """``images``: [B,3,H,W] in [0, 1] (normalised to [-1, 1] as the reference's VaeImageProcessor.preprocess does)
        or already in [-1, 1] (detected by a negative minimum, the same rule diffusers applies)."""
        x = images.to(device=self.device, dtype=torch.float32)
        if float(x.min()) >= 0.0:          # VaeImageProcessor.preprocess (i2sb_pipeline.py:34): [0, 1] inputs -> [-1, 1]
            x = 2.0 * x - 1.0
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
