"""DDIMScheduler restricted to the configuration the reference ships
(/root/reference/configs/ldm/noise_scheduler.json: scaled_linear betas, epsilon prediction,
"leading" spacing with steps_offset 1, no clipping, set_alpha_to_one false) - SURVEY.md 8a-R;
used by afldm/pipelines/ldm_pipeline.py:80-109 and scripts/shift_ldm_ffhq.py:89-104.

The scalar schedule lives on the host (fp32, same operation order as diffusers); the update
``x' = cx * x + ce * eps`` is one kernel.  For eta = 0:
    cx = sqrt(a') / sqrt(a_t),   ce = sqrt(1 - a') - sqrt(a') * sqrt(1 - a_t) / sqrt(a_t).
"""
from __future__ import annotations

import torch

from .. import ops
from ..configs import Config, FFHQ_DDIM


class SchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0195,
                 beta_schedule="scaled_linear", steps_offset=1, set_alpha_to_one=False, clip_sample=False,
                 prediction_type="epsilon", timestep_spacing="leading", **unused):
        if beta_schedule != "scaled_linear" or clip_sample or prediction_type != "epsilon" or \
                timestep_spacing != "leading":
            raise NotImplementedError("only the AF-LDM DDIM configuration is supported")
        self.config = Config(num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
                             beta_schedule=beta_schedule, steps_offset=steps_offset,
                             set_alpha_to_one=set_alpha_to_one, clip_sample=clip_sample,
                             prediction_type=prediction_type, timestep_spacing=timestep_spacing)
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    @classmethod
    def from_pretrained(cls, path_or_repo: str, subfolder="scheduler", **_unused):
        from .. import hub
        return hub.load_scheduler(cls, path_or_repo, subfolder)

    def save_pretrained(self, directory: str):
        from .. import hub
        hub.save_scheduler(self, directory, "DDIMScheduler")

    @classmethod
    def from_config(cls, config=FFHQ_DDIM, **overrides):
        cfg = dict(config)
        cfg.update(overrides)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (torch.arange(0, num_inference_steps) * ratio).flip(0) + self.config.steps_offset
        self.timesteps = ts.to(device) if device is not None else ts

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, timestep: int):
        """(cx, ce) of the eta = 0 update at integer ``timestep``, as fp32 tensors on the host."""
        t = int(timestep)
        prev = t - self.config.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod
        cx = a_p ** 0.5 / a_t ** 0.5
        ce = (1 - a_p) ** 0.5 - a_p ** 0.5 * (1 - a_t) ** 0.5 / a_t ** 0.5
        return cx, ce

    def step(self, model_output, timestep, sample, eta: float = 0.0, return_dict: bool = True, **unused):
        if eta != 0.0:
            raise NotImplementedError("eta > 0 (stochastic DDIM) is not on the AF-LDM path")
        cx, ce = self.coefficients(int(timestep))
        x, e = ops.nhwc(sample), ops.nhwc(model_output)
        prev = ops.nchw_view(ops.axpby(x, e, float(cx), float(ce)))
        return SchedulerOutput(prev) if return_dict else (prev,)
