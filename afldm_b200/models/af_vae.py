"""AutoencoderKL / AliasFreeAutoencoderKL with the diffusers surface the reference scripts touch
(``vae.decode(z).sample``, ``vae.encode(x).latent_dist.{sample,mode,mean}``, ``vae.config``),
forwards running the sm_100a kernels.  Mirrors /root/reference/afldm/models/af_vae.py:8-55
(``AliasFreeAutoencoderKL`` = ``AutoencoderKL`` + ``make_af_vae`` in ``__init__``,
``encode_scale`` / ``decode_scale`` / ``downsample_ratio``) and the diffusers Encoder / Decoder
layout of configs/vae/model_afvae.json."""
from __future__ import annotations

from typing import Sequence

import torch
import torch.nn as nn

from .. import ops
from ..af_modules.af_blocks import act_name
from ..configs import Config, FFHQ_AFVAE
from ..packing import conv_params
from .blocks import DownEncoderBlock2D, UNetMidBlock2D, UpDecoderBlock2D


def _conv(m: nn.Conv2d, x_nchw: torch.Tensor) -> torch.Tensor:
    w, b, k = conv_params(m)
    return ops.nchw_view(ops.conv2d(ops.nhwc(x_nchw), w, b, k))


def _norm_act_conv(norm: nn.GroupNorm, act: nn.Module, conv: nn.Conv2d, x_nchw: torch.Tensor) -> torch.Tensor:
    x = ops.nhwc(x_nchw)
    a = ops.groupnorm_act(x, norm.num_groups, norm.eps, norm.weight, norm.bias, act=act_name(act))
    w, b, k = conv_params(conv)
    return ops.nchw_view(ops.conv2d(a, w, b, k))


class Decoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[-1], 3, padding=1)
        self.mid_block = UNetMidBlock2D(boc[-1], None, boc[-1], 1e-6)      # one head of dim C
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_ch = rev[0]
        for i in range(len(boc)):
            prev, out_ch = out_ch, rev[i]
            self.up_blocks.append(UpDecoderBlock2D(prev, out_ch, layers_per_block + 1, i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, z: torch.Tensor) -> torch.Tensor:
        h = self.mid_block(_conv(self.conv_in, z), None)
        for blk in self.up_blocks:
            h = blk(h)
        return _norm_act_conv(self.conv_norm_out, self.conv_act, self.conv_out, h)


class Encoder(nn.Module):
    def __init__(self, in_channels=3, out_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i in range(len(boc)):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(DownEncoderBlock2D(in_ch, out_ch, layers_per_block, i != len(boc) - 1))
        self.mid_block = UNetMidBlock2D(boc[-1], None, boc[-1], 1e-6)
        self.conv_norm_out = nn.GroupNorm(32, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * out_channels, 3, padding=1)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = _conv(self.conv_in, x)
        for blk in self.down_blocks:
            h = blk(h)
        h = self.mid_block(h, None)
        return _norm_act_conv(self.conv_norm_out, self.conv_act, self.conv_out, h)


class DiagonalGaussianDistribution:
    """diffusers DiagonalGaussianDistribution on the 2*latent_channels encoder moments (tiny tensors;
    plain torch elementwise - not part of the measured path)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype,
                            device=generator.device if generator is not None else "cpu").to(self.mean.device)
        return self.mean + self.std * noise


class _Out:
    def __init__(self, **kw):
        self.__dict__.update(kw)


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block=2, latent_channels=4, scaling_factor=0.6, mid_act=True,
                 down_filtered_act=(False, True, True, True), up_filtered_act=(True, True, True, False),
                 up_rescale=(True, True, True), up_block_types=("UpDecoderBlock2D",) * 4,
                 down_block_types=("DownEncoderBlock2D",) * 4, **unused):
        super().__init__()
        self.config = Config(in_channels=in_channels, out_channels=out_channels,
                             block_out_channels=list(block_out_channels), layers_per_block=layers_per_block,
                             latent_channels=latent_channels, scaling_factor=scaling_factor, mid_act=mid_act,
                             down_filtered_act=list(down_filtered_act), up_filtered_act=list(up_filtered_act),
                             up_rescale=list(up_rescale), up_block_types=list(up_block_types),
                             down_block_types=list(down_block_types))
        # keys of the diffusers config.json that do not change this build's graph are carried through unchanged, so a
        # saved config reloads in diffusers with the same values (sample_size, norm_num_groups, act_fn, use_quant_conv...)
        carried = dict(sample_size=512, norm_num_groups=32, act_fn="silu", use_quant_conv=True, use_post_quant_conv=True,
                       force_upcast=True, latents_mean=None, latents_std=None, shift_factor=None, mid_block_add_attention=True)
        for k, default in carried.items():
            setattr(self.config, k, unused.get(k, default))
        if self.config.norm_num_groups != 32 or self.config.act_fn != "silu":
            raise NotImplementedError("AutoencoderKL: norm_num_groups = 32 and act_fn = 'silu' (the AF-LDM VAE)")
        self.up_block_types = list(up_block_types)        # read by scripts/shift_ldm_ffhq.py:60
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    @classmethod
    def from_config(cls, config=FFHQ_AFVAE, **overrides):
        cfg = dict(config)
        cfg.update(overrides)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    @classmethod
    def from_pretrained(cls, path_or_repo: str, subfolder=None, **_unused):
        """config.json (incl. the alias-free flags of configs/vae/model_afvae.json) + weights of a local directory /
        cached repo.  ``AutoencoderKL`` loads the plain module (scripts then call ``make_af_vae_from_config``);
        ``AliasFreeAutoencoderKL`` applies the surgery in its constructor (af_vae.py:8-55)."""
        from .. import hub
        return hub.load_model(cls, path_or_repo, subfolder)

    def save_pretrained(self, directory: str, safe_serialization: bool = True):
        from .. import hub
        hub.save_model(self, directory, "AutoencoderKL", safe_serialization)

    @property
    def dtype(self):
        return self.post_quant_conv.weight.dtype

    @property
    def device(self):
        return self.post_quant_conv.weight.device

    def encode(self, x: torch.Tensor, return_dict: bool = True):
        moments = _conv(self.quant_conv, self.encoder(x))
        dist = DiagonalGaussianDistribution(ops.to_nchw_contiguous(ops.nhwc(moments)))
        return _Out(latent_dist=dist) if return_dict else (dist,)

    def decode(self, z: torch.Tensor, return_dict: bool = True, **kwargs):
        img = self.decoder(_conv(self.post_quant_conv, z))
        img = ops.to_nchw_contiguous(ops.nhwc(img))
        return _Out(sample=img) if return_dict else (img,)


class AliasFreeAutoencoderKL(AutoencoderKL):
    """af_vae.py:8-55."""

    def __init__(self, *args, mod_mid_act=None, **kwargs):
        super().__init__(*args, **kwargs)
        from ..af_modules.af_api import make_af_vae
        c = self.config
        make_af_vae(self, c.mid_act if mod_mid_act is None else mod_mid_act, c.down_filtered_act,
                    c.up_filtered_act, c.up_rescale)

    @property
    def downsample_ratio(self):
        return 2 ** (len(self.config.block_out_channels) - 1)

    def encode_scale(self, x):
        return self.encode(x).latent_dist.sample() * self.config.scaling_factor

    def decode_scale(self, x):
        return self.decode(x / self.config.scaling_factor).sample
