"""Oracle (test infrastructure): the I2SB scheduler arithmetic in plain PyTorch.

Follows /root/reference/afldm/schedulers/i2sb_scheduler.py: compute_gaussian_product_coef :131-139,
__init__ :188-197 (std_fwd / std_bwd / mu_x0 / mu_x1 / std_sb), set_timesteps :224-302 ("leading"
spacing), previous_timestep :518-531, step :382-459, add_noise :461-483, compute_label :507-513; the
sampling loop that drives it is afldm/pipelines/i2sb_pipeline.py:47-56.  The reference class derives from
diffusers' SchedulerMixin and cannot be imported here (diffusers is absent), so this restatement is
anchored on those lines only: parity unpinned.
"""
from __future__ import annotations

import numpy as np
import torch


def gaussian_product_coef(sigma1, sigma2):
    denom = sigma1 ** 2 + sigma2 ** 2
    return sigma2 ** 2 / denom, sigma1 ** 2 / denom, (sigma1 ** 2 * sigma2 ** 2) / denom


class I2SBScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0195, beta_schedule="scaled_linear",
                 clip_sample=False, clip_sample_range=1.0, timestep_spacing="leading", steps_offset=1, **_ignored):
        assert beta_schedule == "scaled_linear" and timestep_spacing == "leading"
        self.num_train_timesteps, self.steps_offset = num_train_timesteps, steps_offset
        self.clip_sample, self.clip_sample_range = clip_sample, clip_sample_range
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.std_fwd = torch.sqrt(torch.cumsum(self.betas, 0))
        self.std_bwd = torch.sqrt(torch.flip(torch.cumsum(torch.flip(self.betas, dims=[0]), 0), dims=[0]))
        self.mu_x0, self.mu_x1, var = gaussian_product_coef(self.std_fwd, self.std_bwd)
        self.std_sb = torch.sqrt(var)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + self.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def scale_model_input(self, x, t=None):
        return x

    def previous_timestep(self, t):
        n = self.num_inference_steps if self.num_inference_steps else self.num_train_timesteps
        return t - self.num_train_timesteps // n

    def step(self, model_output, timestep, sample, is_ode=False, generator=None, return_dict=True):
        t = int(timestep)
        prev_t = self.previous_timestep(t)
        std_fwd = self.std_fwd.to(sample.device)
        s_t, s_p = std_fwd[t], std_fwd[prev_t]
        s_d = (s_t ** 2 - s_p ** 2).sqrt()
        x0 = sample - s_t * model_output
        if self.clip_sample:
            x0 = x0.clamp(-self.clip_sample_range, self.clip_sample_range)
        mu_x0, mu_xt, var = gaussian_product_coef(s_p, s_d)
        prev = mu_x0 * x0 + mu_xt * sample
        if t > 0 and not is_ode:
            noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype,
                                device=generator.device if generator is not None else "cpu").to(sample.device)
            prev = prev + noise * var.sqrt()
        return (prev,) if not return_dict else type("Out", (), {"prev_sample": prev, "pred_original_sample": x0})()

    def add_noise(self, x0, x1, timesteps, is_ode=False, noise=None):
        e = lambda v: v.to(x0.device)[timesteps].reshape(-1, *([1] * (x0.ndim - 1)))
        xt = e(self.mu_x0) * x0 + e(self.mu_x1) * x1
        if not is_ode:
            xt = xt + e(self.std_sb) * (torch.randn_like(xt) if noise is None else noise)
        return xt

    def compute_label(self, timesteps, x0, xt):
        s = self.std_fwd.to(x0.device)[timesteps].reshape(-1, *([1] * (x0.ndim - 1)))
        return (xt - x0) / s
