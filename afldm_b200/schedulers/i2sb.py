"""I2SB (image-to-image Schroedinger bridge) scheduler of the latent super-resolution sampler, restricted to
the configuration the reference loads it with (configs/ldm/noise_scheduler.json: scaled_linear betas, "leading"
spacing, steps_offset 1, no clipping).  Mirrors /root/reference/afldm/schedulers/i2sb_scheduler.py
(__init__ :188-197, set_timesteps :224-302, previous_timestep :518-531, step :382-459, add_noise :461-483,
compute_label :507-513).

The schedule is scalar host math; because mu_x0 + mu_xt = 1 the update collapses to ONE kernel
    x_prev = x_t - mu_x0 * std_fwd[t] * eps        (+ sqrt(var) * noise when not is_ode).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import ops
from ..configs import Config, FFHQ_DDIM
from .ddim import SchedulerOutput


def compute_gaussian_product_coef(sigma1, sigma2):
    """p1 = N(x_t | x_0, sigma1^2), p2 = N(x_t | x_1, sigma2^2)  ->  coefficients / variance of p1 * p2 (:131-139)."""
    denom = sigma1 ** 2 + sigma2 ** 2
    return sigma2 ** 2 / denom, sigma1 ** 2 / denom, (sigma1 ** 2 * sigma2 ** 2) / denom


class I2SBScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0195, beta_schedule="scaled_linear",
                 clip_sample=False, timestep_spacing="leading", steps_offset=1, **unused):
        if beta_schedule != "scaled_linear" or clip_sample or timestep_spacing != "leading":
            raise NotImplementedError("only the AF-LDM scheduler configuration is supported")
        self.config = Config(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                             beta_schedule=beta_schedule, clip_sample=clip_sample, timestep_spacing=timestep_spacing,
                             steps_offset=steps_offset)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.std_fwd = torch.sqrt(torch.cumsum(self.betas, 0))
        self.std_bwd = torch.sqrt(torch.flip(torch.cumsum(torch.flip(self.betas, dims=[0]), 0), dims=[0]))
        self.mu_x0, self.mu_x1, var = compute_gaussian_product_coef(self.std_fwd, self.std_bwd)
        self.std_sb = torch.sqrt(var)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    @classmethod
    def from_pretrained(cls, path_or_repo: str, subfolder="scheduler", **_unused):
        from .. import hub
        return hub.load_scheduler(cls, path_or_repo, subfolder)

    def save_pretrained(self, directory: str):
        from .. import hub
        hub.save_scheduler(self, directory, "I2SBScheduler")

    @classmethod
    def from_config(cls, config=FFHQ_DDIM, **overrides):
        cfg = dict(config)
        cfg.update(overrides)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    def __len__(self):
        return self.config.num_train_timesteps

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts + self.config.steps_offset).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def previous_timestep(self, timestep: int) -> int:
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def coefficients(self, timestep: int):
        """(c_eps, sqrt_var): x_prev = x_t + c_eps * eps (+ sqrt_var * noise), fp32 host tensors."""
        t = int(timestep)
        prev_t = self.previous_timestep(t)        # negative prev_t indexes from the end, as in the reference (:381 note)
        s_t, s_p = self.std_fwd[t], self.std_fwd[prev_t]
        s_d = (s_t ** 2 - s_p ** 2).sqrt()
        mu_x0, _mu_xt, var = compute_gaussian_product_coef(s_p, s_d)
        return -(mu_x0 * s_t), var.sqrt()

    def step(self, model_output, timestep, sample, is_ode: bool = False, generator=None, return_dict: bool = True):
        t = int(timestep)
        c_eps, sqrt_var = self.coefficients(t)
        x, e = ops.nhwc(sample), ops.nhwc(model_output)
        prev = ops.axpby(x, e, 1.0, float(c_eps))
        if t > 0 and not is_ode:
            gdev = generator.device if generator is not None else torch.device("cpu")
            noise = torch.randn(model_output.shape, generator=generator, device=gdev, dtype=torch.float32).to(prev.device)
            prev = ops.axpby(prev, ops.nhwc(noise), 1.0, float(sqrt_var))
        prev = ops.nchw_view(prev)
        return SchedulerOutput(prev) if return_dict else (prev,)

    def add_noise(self, x0, x1, timesteps, is_ode: bool = False, noise=None):
        """Training-side forward process (:461-483); plain tensor math, not on the inference path."""
        e = lambda v: v.to(x0.device)[timesteps].reshape(-1, *([1] * (x0.ndim - 1)))
        xt = e(self.mu_x0) * x0 + e(self.mu_x1) * x1
        if not is_ode:
            xt = xt + e(self.std_sb) * (torch.randn_like(xt) if noise is None else noise)
        return xt

    def compute_label(self, timesteps, x0, xt):
        s = self.std_fwd.to(x0.device)[timesteps].reshape(-1, *([1] * (x0.ndim - 1)))
        return (xt - x0) / s
