"""Oracle (test infrastructure): the DDIM loops of /root/reference/afldm/pipelines/ldm_pipeline.py on the oracle UNet.

``denoise`` follows :93-109 (scale_model_input -> unet -> scheduler.step, eta = 0), ``ddim_inversion`` follows :133-160
(deterministic DDIM run backwards over ``reversed(scheduler.timesteps)``; ``alpha_prod_t_prev`` of the first step is
``final_alpha_cumprod``)."""
from __future__ import annotations

import torch


@torch.no_grad()
def denoise(unet, scheduler, latents, num_inference_steps=50, on_step=None):
    scheduler.set_timesteps(num_inference_steps)
    for i, t in enumerate(scheduler.timesteps):
        eps = unet(scheduler.scale_model_input(latents, t), t.to(latents.device)).sample
        latents = scheduler.step(eps, int(t), latents, return_dict=False)[0]
        if on_step is not None:
            on_step(i, latents)
    return latents


@torch.no_grad()
def ddim_inversion(unet, scheduler, latent):
    timesteps = list(reversed(scheduler.timesteps.tolist()))
    for i, t in enumerate(timesteps):
        a_t = scheduler.alphas_cumprod[t]
        a_prev = scheduler.alphas_cumprod[timesteps[i - 1]] if i > 0 else scheduler.final_alpha_cumprod
        mu, mu_prev = a_t ** 0.5, a_prev ** 0.5
        sigma, sigma_prev = (1 - a_t) ** 0.5, (1 - a_prev) ** 0.5
        eps = unet(latent, torch.tensor(t, device=latent.device)).sample
        pred_x0 = (latent - sigma_prev * eps) / mu_prev
        latent = mu * pred_x0 + sigma * eps
    return latent
