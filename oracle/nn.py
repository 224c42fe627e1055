"""Oracle (test infrastructure): plain-PyTorch restatement of the diffusers 0.32.1 blocks that
the reference mutates, with diffusers-identical parameter names.

diffusers is a third-party dependency that is absent from /root/reference and from this image
(requirements.txt:1 un-pinned; configs/ldm/model_unet.json:3 records 0.32.1), so this module
follows the published algorithm of that release.  PARITY UNPINNED at this boundary: there are
no reference tests or golden vectors for it.  Call sites in the reference:
afldm/pipelines/ldm_pipeline.py:1-4,93,103-109, afldm/af_modules/af_api.py:1-3,
afldm/af_modules/af_blocks.py:6-7,45,109, afldm/models/af_vae.py:2,8,
afldm/pipelines/cross_frame_attn.py:3,54.
"""
from __future__ import annotations

import math
from types import SimpleNamespace
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- embeddings
def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True,
                       freq_shift: float = 0.0, max_period: float = 10000.0) -> torch.Tensor:
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32, device=t.device)
    freqs = torch.exp(-math.log(max_period) * k / (half - freq_shift))
    arg = t[:, None].float() * freqs[None, :]
    s, c = torch.sin(arg), torch.cos(arg)
    return torch.cat([c, s], dim=-1) if flip_sin_to_cos else torch.cat([s, c], dim=-1)


class TimestepEmbedding(nn.Module):
    def __init__(self, in_dim: int, dim: int):
        super().__init__()
        self.linear_1 = nn.Linear(in_dim, dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(dim, dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


# ----------------------------------------------------------------------------- resnet
class ResnetBlock2D(nn.Module):
    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int],
                 groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        self.nonlinearity = nn.SiLU()          # single shared activation: surgery replaces it
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb=None):
        h = self.conv1(self.nonlinearity(self.norm1(x)))
        if self.time_emb_proj is not None and temb is not None:
            h = h + self.time_emb_proj(self.nonlinearity(temb))[:, :, None, None]
        h = self.conv2(self.nonlinearity(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


# ----------------------------------------------------------------------------- attention
class AttnProcessor2_0:
    """diffusers AttnProcessor2_0 for the legacy attention-block form (4-D input)."""

    def __call__(self, attn: "Attention", hidden_states, encoder_hidden_states=None,
                 attention_mask=None, temb=None):
        residual = hidden_states
        b, c, hh, ww = hidden_states.shape
        x = hidden_states.view(b, c, hh * ww).transpose(1, 2)
        if attn.group_norm is not None:
            x = attn.group_norm(x.transpose(1, 2)).transpose(1, 2)
        q = attn.to_q(x)
        src = x if encoder_hidden_states is None else encoder_hidden_states
        k, v = attn.to_k(src), attn.to_v(src)
        d = k.shape[-1] // attn.heads
        q = q.view(b, -1, attn.heads, d).transpose(1, 2)
        k = k.view(b, -1, attn.heads, d).transpose(1, 2)
        v = v.view(b, -1, attn.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * d)
        o = attn.to_out[0](o)
        o = o.transpose(-1, -2).reshape(b, c, hh, ww)
        if attn.residual_connection:
            o = o + residual
        return o / attn.rescale_output_factor


class Attention(nn.Module):
    def __init__(self, channels: int, heads: int, dim_head: int, groups: int = 32, eps: float = 1e-5):
        super().__init__()
        assert heads * dim_head == channels
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        self.residual_connection = True
        self.rescale_output_factor = 1.0
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.to_q = nn.Linear(channels, channels)
        self.to_k = nn.Linear(channels, channels)
        self.to_v = nn.Linear(channels, channels)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def set_processor(self, p):
        self.processor = p

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, **kw)


# ----------------------------------------------------------------------------- resamplers
class Downsample2D(nn.Module):
    def __init__(self, channels, use_conv=True, out_channels=None, padding=1, name="op"):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.padding, self.name = use_conv, padding, name
        self.norm = None
        self.conv = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, x, *a, **k):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1))
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels, use_conv=True, out_channels=None, name="conv"):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.name = use_conv, name
        self.norm = None
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)

    def forward(self, x, output_size=None, *a, **k):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


# ----------------------------------------------------------------------------- UNet blocks
class DownBlock(nn.Module):
    """AttnDownBlock2D / DownBlock2D (attention optional)."""

    def __init__(self, cin, cout, temb, layers, add_down, attn_head_dim, eps, down_padding=1):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, temb, eps=eps) for i in range(layers)])
        self.attentions = nn.ModuleList(
            [Attention(cout, cout // attn_head_dim, attn_head_dim, eps=eps) for _ in range(layers)]
        ) if attn_head_dim else None
        self.downsamplers = nn.ModuleList(
            [Downsample2D(cout, True, cout, padding=down_padding, name="op")]) if add_down else None

    def forward(self, h, temb=None):
        outs = ()
        for i, r in enumerate(self.resnets):
            h = r(h, temb)
            if self.attentions is not None:
                h = self.attentions[i](h)
            outs += (h,)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs += (h,)
        return h, outs


class MidBlock(nn.Module):
    def __init__(self, c, temb, attn_head_dim, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, eps=eps), ResnetBlock2D(c, c, temb, eps=eps)])
        self.attentions = nn.ModuleList([Attention(c, c // attn_head_dim, attn_head_dim, eps=eps)])

    def forward(self, h, temb=None):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h)
        return self.resnets[1](h, temb)


class UpBlock(nn.Module):
    """AttnUpBlock2D / UpBlock2D."""

    def __init__(self, cin, cout, prev, temb, layers, add_up, attn_head_dim, eps):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            first = prev if i == 0 else cout
            rs.append(ResnetBlock2D(first + skip, cout, temb, eps=eps))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList(
            [Attention(cout, cout // attn_head_dim, attn_head_dim, eps=eps) for _ in range(layers)]
        ) if attn_head_dim else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout, True, cout)]) if add_up else None

    def forward(self, h, skips, temb=None):
        skips = list(skips)
        for i, r in enumerate(self.resnets):
            h = r(torch.cat([h, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                h = self.attentions[i](h)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class UNet2DModel(nn.Module):
    """diffusers UNet2DModel for the block types in configs/ldm/model_unet.json."""

    def __init__(self, sample_size=32, in_channels=4, out_channels=4,
                 block_out_channels: Sequence[int] = (192, 384, 384, 768, 768),
                 down_block_types=("AttnDownBlock2D",) * 4 + ("DownBlock2D",),
                 up_block_types=("UpBlock2D",) + ("AttnUpBlock2D",) * 4,
                 layers_per_block=2, attention_head_dim=24, norm_eps=1e-5, **_ignored):
        super().__init__()
        boc = list(block_out_channels)
        self.config = SimpleNamespace(sample_size=sample_size, in_channels=in_channels,
                                      out_channels=out_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block,
                                      attention_head_dim=attention_head_dim)
        temb = boc[0] * 4
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, ty in enumerate(down_block_types):
            cin, cout = cout, boc[i]
            self.down_blocks.append(DownBlock(cin, cout, temb, layers_per_block, i != len(boc) - 1,
                                              attention_head_dim if ty.startswith("Attn") else None, norm_eps))
        self.mid_block = MidBlock(boc[-1], temb, attention_head_dim, norm_eps)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        cout = rev[0]
        for i, ty in enumerate(up_block_types):
            prev, cout = cout, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(UpBlock(cin, cout, prev, temb, layers_per_block + 1, i != len(boc) - 1,
                                          attention_head_dim if ty.startswith("Attn") else None, norm_eps))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timestep, return_dict=True):
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.long, device=sample.device)
        elif t.ndim == 0:
            t = t[None].to(sample.device)
        t = t * torch.ones(sample.shape[0], dtype=t.dtype, device=t.device)
        emb = self.time_embedding(timestep_embedding(t, self.config.block_out_channels[0]).to(sample.dtype))
        h = self.conv_in(sample)
        skips = (h,)
        for blk in self.down_blocks:
            h, outs = blk(h, emb)
            skips += outs
        h = self.mid_block(h, emb)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            h = blk(h, skips[-n:], emb)
            skips = skips[:-n]
        h = self.conv_out(self.conv_act(self.conv_norm_out(h)))
        if return_dict:
            return SimpleNamespace(sample=h)
        return (h,)


# ----------------------------------------------------------------------------- VAE
class UpDecoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, add_up, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, None, eps=eps) for i in range(layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(cout, True, cout)]) if add_up else None

    def forward(self, h, temb=None):
        for r in self.resnets:
            h = r(h, None)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class DownEncoderBlock(nn.Module):
    def __init__(self, cin, cout, layers, add_down, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList(
            [ResnetBlock2D(cin if i == 0 else cout, cout, None, eps=eps) for i in range(layers)])
        self.downsamplers = nn.ModuleList(
            [Downsample2D(cout, True, cout, padding=0, name="op")]) if add_down else None

    def forward(self, h):
        for r in self.resnets:
            h = r(h, None)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
        return h


class Decoder(nn.Module):
    def __init__(self, in_channels=4, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[-1], 3, padding=1)
        self.mid_block = MidBlock(boc[-1], None, boc[-1], 1e-6)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        cout = rev[0]
        for i in range(len(boc)):
            prev, cout = cout, rev[i]
            self.up_blocks.append(UpDecoderBlock(prev, cout, layers_per_block + 1, i != len(boc) - 1))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    def forward(self, z):
        h = self.mid_block(self.conv_in(z), None)
        for blk in self.up_blocks:
            h = blk(h)
        return self.conv_out(self.conv_act(self.conv_norm_out(h)))


class Encoder(nn.Module):
    def __init__(self, in_channels=3, out_channels=4, block_out_channels=(128, 256, 512, 512), layers_per_block=2):
        super().__init__()
        boc = list(block_out_channels)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i in range(len(boc)):
            cin, cout = cout, boc[i]
            self.down_blocks.append(DownEncoderBlock(cin, cout, layers_per_block, i != len(boc) - 1))
        self.mid_block = MidBlock(boc[-1], None, boc[-1], 1e-6)
        self.conv_norm_out = nn.GroupNorm(32, boc[-1], eps=1e-6)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[-1], 2 * out_channels, 3, padding=1)

    def forward(self, x):
        h = self.conv_in(x)
        for blk in self.down_blocks:
            h = blk(h)
        h = self.mid_block(h, None)
        return self.conv_out(self.conv_act(self.conv_norm_out(h)))


class DiagonalGaussian:
    def __init__(self, params):
        self.mean, self.logvar = torch.chunk(params, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype).to(self.mean.device)
        return self.mean + self.std * noise


class AutoencoderKL(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512),
                 layers_per_block=2, latent_channels=4, scaling_factor=0.6,
                 mid_act=True, down_filtered_act=(False, True, True, True),
                 up_filtered_act=(True, True, True, False), up_rescale=(True, True, True), **_ignored):
        super().__init__()
        self.config = SimpleNamespace(scaling_factor=scaling_factor, block_out_channels=list(block_out_channels),
                                      latent_channels=latent_channels, mid_act=mid_act,
                                      down_filtered_act=list(down_filtered_act),
                                      up_filtered_act=list(up_filtered_act), up_rescale=list(up_rescale))
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block)
        self.decoder = Decoder(latent_channels, out_channels, block_out_channels, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.post_quant_conv = nn.Conv2d(latent_channels, latent_channels, 1)

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussian(self.quant_conv(self.encoder(x))))

    def decode(self, z, return_dict=True):
        img = self.decoder(self.post_quant_conv(z))
        return SimpleNamespace(sample=img) if return_dict else (img,)


# ----------------------------------------------------------------------------- DDIM
class DDIMScheduler:
    """diffusers DDIMScheduler restricted to configs/ldm/noise_scheduler.json
    (scaled_linear betas, epsilon prediction, leading spacing, steps_offset, no clipping)."""

    def __init__(self, num_train_timesteps=1000, beta_start=0.0015, beta_end=0.0195,
                 steps_offset=1, set_alpha_to_one=False, **_ignored):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                    dtype=torch.float32) ** 2
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    def set_timesteps(self, n, device=None):
        self.num_inference_steps = n
        ratio = self.num_train_timesteps // n
        ts = (torch.arange(0, n) * ratio).flip(0) + self.steps_offset
        self.timesteps = ts.to(device) if device is not None else ts

    def scale_model_input(self, x, t=None):
        return x

    def step(self, eps, t, x, eta=0.0, return_dict=True, **_kw):
        t = int(t)
        tp = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[tp] if tp >= 0 else self.final_alpha_cumprod
        x0 = (x - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        prev = a_p ** 0.5 * x0 + (1 - a_p) ** 0.5 * eps      # eta = 0: variance term vanishes
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0) if return_dict else (prev,)
