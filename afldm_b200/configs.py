"""Model / scheduler hyper-parameters of the released AF-LDM FFHQ models, restated from the
reference's JSON files (/root/reference/configs/ldm/model_unet.json,
configs/vae/model_afvae.json, configs/ldm/noise_scheduler.json) so that nothing has to be read from
the reference tree at run time.  ``Config`` mimics diffusers' FrozenDict (attribute + item access).
"""
from __future__ import annotations

import json


class Config(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def load_json(path: str) -> Config:
    with open(path) as f:
        return Config(json.load(f))


FFHQ_UNET = Config(
    sample_size=32, in_channels=4, out_channels=4, act_fn="silu", add_attention=True,
    attention_head_dim=24, block_out_channels=[192, 384, 384, 768, 768],
    down_block_types=["AttnDownBlock2D"] * 4 + ["DownBlock2D"],
    up_block_types=["UpBlock2D"] + ["AttnUpBlock2D"] * 4,
    downsample_padding=1, flip_sin_to_cos=True, freq_shift=0, layers_per_block=2,
    norm_eps=1e-5, norm_num_groups=32, time_embedding_type="positional")

FFHQ_AFVAE = Config(
    in_channels=3, out_channels=3, latent_channels=4, act_fn="silu",
    block_out_channels=[128, 256, 512, 512], layers_per_block=2, norm_num_groups=32,
    down_block_types=["DownEncoderBlock2D"] * 4, up_block_types=["UpDecoderBlock2D"] * 4,
    sample_size=512, scaling_factor=0.6, mid_act=True,
    down_filtered_act=[False, True, True, True], up_filtered_act=[True, True, True, False],
    up_rescale=[True, True, True], use_quant_conv=True, use_post_quant_conv=True,
    mid_block_add_attention=True)

FFHQ_DDIM = Config(
    num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0195, beta_schedule="scaled_linear",
    clip_sample=False, prediction_type="epsilon", set_alpha_to_one=False, steps_offset=1,
    timestep_spacing="leading")
