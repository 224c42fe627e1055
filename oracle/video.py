"""Oracle (test infrastructure): the frame loop of /root/reference/afldm/pipelines/video_equiv_editing_pipeline.py on the
oracle models - ``image2latent`` :214-226, ``ddim_inversion`` :174-211 (no guidance, ``attn_invert=True``),
``get_timesteps`` :319-327, the STORE pass ``save_activations`` :612-649 and the per-frame LOAD denoising loop :657-697
with classifier-free guidance, one frame (batch 2) at a time exactly as the reference does."""
from __future__ import annotations

import torch

from . import cross_frame as OC
from . import nn_cond


@torch.no_grad()
def ddim_inversion(unet, scheduler, state, latent, timesteps, cond):
    ts = list(reversed([int(t) for t in timesteps]))
    for i, t in enumerate(ts):
        a_t = scheduler.alphas_cumprod[t]
        a_prev = scheduler.alphas_cumprod[ts[i - 1]] if i > 0 else scheduler.final_alpha_cumprod
        mu, mu_prev = a_t ** 0.5, a_prev ** 0.5
        sigma, sigma_prev = (1 - a_t) ** 0.5, (1 - a_prev) ** 0.5
        state.set_timestep(t)
        eps = unet(latent, torch.tensor(t, device=latent.device), encoder_hidden_states=cond.repeat(latent.shape[0], 1, 1)).sample
        pred_x0 = (latent - sigma_prev * eps) / mu_prev
        latent = mu * pred_x0 + sigma * eps
    return latent


@torch.no_grad()
def video_edit(unet, vae, scheduler, frames, pos, neg, inv, num_inference_steps, strength, guidance_scale, decode=True):
    """frames: list of [1,3,H,W] tensors in [-1,1]; pos / neg / inv: [1,77,D].  Returns (latents [F,4,h,w], images | None)."""
    dev = frames[0].device
    state = OC.AttnState()
    OC.set_attn_processor(unet, lambda: OC.CrossFrameAttnProcessor(state, base=nn_cond.AttnProcessor2_0()))
    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    if strength >= 0:
        init = min(int(num_inference_steps * strength), num_inference_steps)
        timesteps = timesteps[max(num_inference_steps - init, 0):]
    ts = [int(t) for t in timesteps]
    sf = vae.scaling_factor if hasattr(vae, "scaling_factor") else vae.config.scaling_factor

    def image2latent(img):
        return vae.encode(img).latent_dist.mean * sf

    state.reset()
    lat = [ddim_inversion(unet, scheduler, state, image2latent(frames[0]), ts, inv)]
    state.to_load()
    for f in frames[1:]:
        lat.append(ddim_inversion(unet, scheduler, state, image2latent(f), ts, inv))
    latents = torch.cat(lat, dim=0)
    ehs = torch.cat([neg, pos], dim=0)

    def cfg_eps(x, t):
        e = unet(torch.cat([x] * 2), torch.tensor(t, device=dev), encoder_hidden_states=ehs).sample
        eu, ec = e.chunk(2)
        return eu + guidance_scale * (ec - eu)

    state.reset()
    state.set_store_id(0)
    x = latents[:1]
    for t in ts:
        state.set_timestep(t)
        x = scheduler.step(cfg_eps(x, t), t, x, return_dict=False)[0]
    state.to_load()
    for t in ts:
        state.set_timestep(t)
        eps = torch.cat([cfg_eps(latents[i:i + 1], t) for i in range(latents.shape[0])])
        latents = scheduler.step(eps, t, latents, return_dict=False)[0]
    images = torch.cat([vae.decode(latents[i:i + 1] / sf).sample for i in range(latents.shape[0])]) if decode else None
    return latents, images
