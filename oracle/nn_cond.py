"""Oracle (test infrastructure): plain-PyTorch restatement of diffusers 0.32.1 ``UNet2DConditionModel`` in its SD-1.5
configuration - the model /root/reference/afldm/pipelines/video_equiv_editing_pipeline.py:196-197, :636-641, :680-686
evaluates (``StableDiffusionPipeline`` components; ``make_af_unet`` applied by scripts/video_editing.py).

Third-party arithmetic, absent from /root/reference and from this image: follows the published algorithm of that
release (``CrossAttnDownBlock2D`` / ``UNetMidBlock2DCrossAttn`` / ``CrossAttnUpBlock2D``, ``Transformer2DModel`` with
``use_linear_projection=False``, ``BasicTransformerBlock`` = LayerNorm -> attn1 -> LayerNorm -> attn2 -> LayerNorm ->
GEGLU feed-forward, ``attention_head_dim: 8`` meaning 8 HEADS).  PARITY UNPINNED at this boundary, like oracle/nn.py;
parameter names are diffusers' so a released state_dict would load."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .nn import Downsample2D, ResnetBlock2D, TimestepEmbedding, Upsample2D, timestep_embedding


class AttnProcessor2_0:
    """diffusers AttnProcessor2_0 for 3-D input [B, N, C] (transformer blocks)."""

    def __call__(self, attn, hidden_states, encoder_hidden_states=None, attention_mask=None, temb=None):
        b = hidden_states.shape[0]
        q = attn.to_q(hidden_states)
        src = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        k, v = attn.to_k(src), attn.to_v(src)
        d = k.shape[-1] // attn.heads
        q = q.view(b, -1, attn.heads, d).transpose(1, 2)
        k = k.view(src.shape[0], -1, attn.heads, d).transpose(1, 2)
        v = v.view(src.shape[0], -1, attn.heads, d).transpose(1, 2)
        o = F.scaled_dot_product_attention(q, k, v)
        o = o.transpose(1, 2).reshape(b, -1, attn.heads * d)
        return attn.to_out[0](o)


class Attention(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim=None):
        super().__init__()
        self.heads, self.dim_head = heads, dim_head
        self.group_norm = None
        self.residual_connection = False
        self.rescale_output_factor = 1.0
        inner = heads * dim_head
        self.to_q = nn.Linear(dim, inner, bias=False)
        self.to_k = nn.Linear(cross_attention_dim or dim, inner, bias=False)
        self.to_v = nn.Linear(cross_attention_dim or dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, dim), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def set_processor(self, p):
        self.processor = p

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, **kw)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)

    def forward(self, h, ehs):
        h = self.attn1(self.norm1(h), encoder_hidden_states=None) + h
        h = self.attn2(self.norm2(h), encoder_hidden_states=ehs) + h
        return self.ff(self.norm3(h)) + h


class Transformer2DModel(nn.Module):
    def __init__(self, heads, dim_head, in_channels, cross_attention_dim):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, x, ehs):
        b, c, hh, ww = x.shape
        h = self.proj_in(self.norm(x))
        inner = h.shape[1]
        h = h.permute(0, 2, 3, 1).reshape(b, hh * ww, inner)
        for blk in self.transformer_blocks:
            h = blk(h, ehs)
        h = h.reshape(b, hh, ww, inner).permute(0, 3, 1, 2).contiguous()
        return self.proj_out(h) + x


class CrossAttnDownBlock(nn.Module):
    def __init__(self, cin, cout, temb, layers, add_down, heads, cross_dim, eps, attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(cin if i == 0 else cout, cout, temb, eps=eps) for i in range(layers)])
        self.attentions = nn.ModuleList(
            [Transformer2DModel(heads, cout // heads, cout, cross_dim) for _ in range(layers)]) if attention else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout, True, cout, padding=1, name="op")]) if add_down else None

    def forward(self, h, temb, ehs):
        outs = ()
        for i, r in enumerate(self.resnets):
            h = r(h, temb)
            if self.attentions is not None:
                h = self.attentions[i](h, ehs)
            outs += (h,)
        if self.downsamplers is not None:
            h = self.downsamplers[0](h)
            outs += (h,)
        return h, outs


class MidBlockCrossAttn(nn.Module):
    def __init__(self, c, temb, heads, cross_dim, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(c, c, temb, eps=eps), ResnetBlock2D(c, c, temb, eps=eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, c // heads, c, cross_dim)])

    def forward(self, h, temb, ehs):
        h = self.resnets[0](h, temb)
        h = self.attentions[0](h, ehs)
        return self.resnets[1](h, temb)


class CrossAttnUpBlock(nn.Module):
    def __init__(self, cin, cout, prev, temb, layers, add_up, heads, cross_dim, eps, attention=True):
        super().__init__()
        rs = []
        for i in range(layers):
            skip = cin if i == layers - 1 else cout
            first = prev if i == 0 else cout
            rs.append(ResnetBlock2D(first + skip, cout, temb, eps=eps))
        self.resnets = nn.ModuleList(rs)
        self.attentions = nn.ModuleList(
            [Transformer2DModel(heads, cout // heads, cout, cross_dim) for _ in range(layers)]) if attention else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout, True, cout)]) if add_up else None

    def forward(self, h, skips, temb, ehs):
        skips = list(skips)
        for i, r in enumerate(self.resnets):
            h = r(torch.cat([h, skips.pop()], dim=1), temb)
            if self.attentions is not None:
                h = self.attentions[i](h, ehs)
        if self.upsamplers is not None:
            h = self.upsamplers[0](h)
        return h


class UNet2DConditionModel(nn.Module):
    def __init__(self, sample_size=64, in_channels=4, out_channels=4,
                 block_out_channels: Sequence[int] = (320, 640, 1280, 1280),
                 down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                 up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
                 layers_per_block=2, attention_head_dim=8, cross_attention_dim=768, norm_eps=1e-5, **_ignored):
        super().__init__()
        boc = list(block_out_channels)
        heads = attention_head_dim
        self.config = SimpleNamespace(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                                      block_out_channels=boc, layers_per_block=layers_per_block,
                                      attention_head_dim=attention_head_dim, cross_attention_dim=cross_attention_dim)
        temb = boc[0] * 4
        self.time_embedding = TimestepEmbedding(boc[0], temb)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = boc[0]
        for i, ty in enumerate(down_block_types):
            cin, cout = cout, boc[i]
            self.down_blocks.append(CrossAttnDownBlock(cin, cout, temb, layers_per_block, i != len(boc) - 1, heads,
                                                       cross_attention_dim, norm_eps, ty.startswith("CrossAttn")))
        self.mid_block = MidBlockCrossAttn(boc[-1], temb, heads, cross_attention_dim, norm_eps)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        cout = rev[0]
        for i, ty in enumerate(up_block_types):
            prev, cout = cout, rev[i]
            cin = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(CrossAttnUpBlock(cin, cout, prev, temb, layers_per_block + 1, i != len(boc) - 1, heads,
                                                   cross_attention_dim, norm_eps, ty.startswith("CrossAttn")))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    def forward(self, sample, timestep, encoder_hidden_states=None, return_dict=True):
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
            h, outs = blk(h, emb, encoder_hidden_states)
            skips += outs
        h = self.mid_block(h, emb, encoder_hidden_states)
        for blk in self.up_blocks:
            n = len(blk.resnets)
            h = blk(h, skips[-n:], emb, encoder_hidden_states)
            skips = skips[:-n]
        h = self.conv_out(self.conv_act(self.conv_norm_out(h)))
        return SimpleNamespace(sample=h) if return_dict else (h,)
