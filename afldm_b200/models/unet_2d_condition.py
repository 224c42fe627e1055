"""SD-1.5 ``UNet2DConditionModel`` with the diffusers surface the reference's video pipeline touches
(``unet(sample, t, encoder_hidden_states=...)``, ``.config``, ``.attn_processors`` / ``.set_attn_processor``,
``down_blocks / mid_block / up_blocks`` for ``make_af_unet``), forwards running the sm_100a kernels.

Call sites in the reference: afldm/pipelines/video_equiv_editing_pipeline.py:196-197 (DDIM inversion), :636-641
(activation store pass), :680-686 (per-frame denoising with classifier-free guidance); the model itself is diffusers'
(0.32.1) - restated from that release like the rest of SURVEY.md 8a-R: ``CrossAttnDownBlock2D`` x3 + ``DownBlock2D``,
``UNetMidBlock2DCrossAttn``, ``UpBlock2D`` + ``CrossAttnUpBlock2D`` x3; every attention is a ``Transformer2DModel``
(GroupNorm(32, eps 1e-6) -> 1x1 conv -> ``BasicTransformerBlock`` -> 1x1 conv, + residual) whose block is
LayerNorm -> self-attention ``attn1`` -> LayerNorm -> cross-attention ``attn2`` on the 77 x 768 text states ->
LayerNorm -> GEGLU feed-forward, each with a residual; 8 heads of 40 / 80 / 160 channels, bias-free q / k / v.
Parameter names are the diffusers ones, so an SD-1.5 ``state_dict`` loads unchanged.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Union

import torch
import torch.nn as nn

from .. import ops
from ..af_modules.af_blocks import act_name
from ..configs import Config
from ..packing import conv_params, conv_params_padded, fused_linear_params
from .blocks import Attention, Downsample2D, ResnetBlock2D, Upsample2D
from .unet_2d import TimestepEmbedding, UNet2DOutput

SD15_UNET = Config(
    sample_size=64, in_channels=4, out_channels=4, act_fn="silu", attention_head_dim=8,
    block_out_channels=[320, 640, 1280, 1280], cross_attention_dim=768,
    down_block_types=["CrossAttnDownBlock2D"] * 3 + ["DownBlock2D"],
    up_block_types=["UpBlock2D"] + ["CrossAttnUpBlock2D"] * 3, mid_block_type="UNetMidBlock2DCrossAttn",
    layers_per_block=2, downsample_padding=1, flip_sin_to_cos=True, freq_shift=0, norm_eps=1e-5, norm_num_groups=32,
    center_input_sample=False, mid_block_scale_factor=1, use_linear_projection=False, only_cross_attention=False)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)


class FeedForward(nn.Module):
    """diffusers FeedForward(activation_fn='geglu'): ``net = [GEGLU, Dropout, Linear]``."""

    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim)])

    def forward(self, x: torch.Tensor, residual: torch.Tensor) -> torch.Tensor:
        """x, residual [B, N, C] -> ff(x) + residual (the residual add rides in the second projection's epilogue)."""
        b, n, c = x.shape
        w, bias, k = conv_params(self.net[0].proj)
        proj = ops.conv2d(x.view(b, n, 1, c), w, bias, k)                       # [B, N, 1, 8C]
        g = ops.geglu(proj.view(b, n, proj.shape[-1]))                          # [B, N, 4C]
        w2, b2, k2 = conv_params(self.net[2])
        out = ops.conv2d(g.view(b, n, 1, g.shape[-1]), w2, b2, k2, residual=residual.view(b, n, 1, c))
        return out.view(b, n, c)


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head, bias=False, norm=False, residual=False)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim=cross_attention_dim, bias=False, norm=False,
                               residual=False)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    def forward(self, h: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor]) -> torch.Tensor:
        """h [B, N, C] (contiguous)."""
        n1 = ops.layer_norm(h, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        h = ops.axpby(self.attn1(n1, encoder_hidden_states=None).contiguous(), h, 1.0, 1.0)
        n2 = ops.layer_norm(h, self.norm2.weight, self.norm2.bias, self.norm2.eps)
        h = ops.axpby(self.attn2(n2, encoder_hidden_states=encoder_hidden_states).contiguous(), h, 1.0, 1.0)
        n3 = ops.layer_norm(h, self.norm3.weight, self.norm3.bias, self.norm3.eps)
        return self.ff(n3, h)


class Transformer2DModel(nn.Module):
    """diffusers Transformer2DModel, continuous input, conv projections (``use_linear_projection=False``), one block."""

    def __init__(self, heads: int, dim_head: int, in_channels: int, cross_attention_dim: int, groups: int = 32):
        super().__init__()
        inner = heads * dim_head
        self.norm = nn.GroupNorm(groups, in_channels, eps=1e-6, affine=True)
        self.proj_in = nn.Conv2d(in_channels, inner, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.proj_out = nn.Conv2d(inner, in_channels, 1)

    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None) -> torch.Tensor:
        x = ops.nhwc(hidden_states)
        b, hh, ww, c = x.shape
        gn = self.norm
        xn = ops.groupnorm_act(x, gn.num_groups, gn.eps, gn.weight, gn.bias, act="identity")
        w, bias, k = conv_params(self.proj_in)
        h = ops.conv2d(xn, w, bias, k).view(b, hh * ww, -1)                     # tokens [B, HW, inner]: NHWC flattened
        for blk in self.transformer_blocks:
            h = blk(h, encoder_hidden_states)
        w, bias, k = conv_params(self.proj_out)
        out = ops.conv2d(h.view(b, hh, ww, h.shape[-1]), w, bias, k, residual=x, gn_stats=True)
        return ops.nchw_view(out)


class CrossAttnDownBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample, heads, cross_attention_dim,
                 eps, downsample_padding=1, attention=True):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps=eps)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim)
            for _ in range(num_layers)]) if attention else None
        self.downsamplers = nn.ModuleList(
            [Downsample2D(out_channels, True, out_channels, padding=downsample_padding, name="op")]) if add_downsample else None

    def forward(self, hidden_states, temb, temb_projs, encoder_hidden_states):
        outputs = ()
        for i, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb, temb_projs[i])
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states)
            outputs += (hidden_states,)
        if self.downsamplers is not None:
            hidden_states = self.downsamplers[0](hidden_states)
            outputs += (hidden_states,)
        return hidden_states, outputs


class UNetMidBlock2DCrossAttn(nn.Module):
    def __init__(self, channels, temb_channels, heads, cross_attention_dim, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, temb_channels, eps=eps),
                                      ResnetBlock2D(channels, channels, temb_channels, eps=eps)])
        self.attentions = nn.ModuleList([Transformer2DModel(heads, channels // heads, channels, cross_attention_dim)])

    def forward(self, hidden_states, temb, temb_projs, encoder_hidden_states):
        hidden_states = self.resnets[0](hidden_states, temb, temb_projs[0])
        hidden_states = self.attentions[0](hidden_states, encoder_hidden_states)
        return self.resnets[1](hidden_states, temb, temb_projs[1])


class CrossAttnUpBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers, add_upsample, heads,
                 cross_attention_dim, eps, attention=True):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip_channels = in_channels if i == num_layers - 1 else out_channels
            first = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(first + skip_channels, out_channels, temb_channels, eps=eps))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = nn.ModuleList([
            Transformer2DModel(heads, out_channels // heads, out_channels, cross_attention_dim)
            for _ in range(num_layers)]) if attention else None
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, temb_projs, encoder_hidden_states):
        skips = list(res_hidden_states_tuple)
        for i, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb, temb_projs[i], skip=skips.pop())
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states, encoder_hidden_states)
        if self.upsamplers is not None:
            hidden_states = self.upsamplers[0](hidden_states)
        return hidden_states


class UNet2DConditionModel(nn.Module):
    def __init__(self, sample_size=64, in_channels=4, out_channels=4,
                 block_out_channels: Sequence[int] = (320, 640, 1280, 1280),
                 down_block_types=("CrossAttnDownBlock2D",) * 3 + ("DownBlock2D",),
                 up_block_types=("UpBlock2D",) + ("CrossAttnUpBlock2D",) * 3,
                 layers_per_block=2, attention_head_dim=8, cross_attention_dim=768, norm_eps=1e-5, norm_num_groups=32,
                 downsample_padding=1, flip_sin_to_cos=True, freq_shift=0, act_fn="silu", **unused):
        super().__init__()
        if not flip_sin_to_cos or freq_shift != 0 or act_fn != "silu" or norm_num_groups != 32:
            raise NotImplementedError("only the SD-1.5 UNet configuration family is supported")
        if unused.get("use_linear_projection") or unused.get("only_cross_attention"):
            raise NotImplementedError("use_linear_projection / only_cross_attention are not SD-1.5 settings")
        boc = list(block_out_channels)
        heads = attention_head_dim          # SD-1.5's config stores the NUMBER OF HEADS under this key (diffusers quirk)
        self.config = Config(sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
                             block_out_channels=boc, down_block_types=list(down_block_types),
                             up_block_types=list(up_block_types), layers_per_block=layers_per_block,
                             attention_head_dim=attention_head_dim, cross_attention_dim=cross_attention_dim,
                             norm_eps=norm_eps, norm_num_groups=norm_num_groups, downsample_padding=downsample_padding,
                             flip_sin_to_cos=flip_sin_to_cos, freq_shift=freq_shift, act_fn=act_fn)
        temb_ch = boc[0] * 4
        self.time_embedding = TimestepEmbedding(boc[0], temb_ch)
        self.conv_in = nn.Conv2d(in_channels, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, kind in enumerate(down_block_types):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(CrossAttnDownBlock2D(
                in_ch, out_ch, temb_ch, layers_per_block, i != len(boc) - 1, heads, cross_attention_dim, norm_eps,
                downsample_padding, attention=kind.startswith("CrossAttn")))
        self.mid_block = UNetMidBlock2DCrossAttn(boc[-1], temb_ch, heads, cross_attention_dim, norm_eps)
        self.up_blocks = nn.ModuleList()
        rev = boc[::-1]
        out_ch = rev[0]
        for i, kind in enumerate(up_block_types):
            prev, out_ch = out_ch, rev[i]
            in_ch = rev[min(i + 1, len(boc) - 1)]
            self.up_blocks.append(CrossAttnUpBlock2D(
                in_ch, out_ch, prev, temb_ch, layers_per_block + 1, i != len(boc) - 1, heads, cross_attention_dim,
                norm_eps, attention=kind.startswith("CrossAttn")))
        self.conv_norm_out = nn.GroupNorm(norm_num_groups, boc[0], eps=norm_eps)
        self.conv_act = nn.SiLU()
        self.conv_out = nn.Conv2d(boc[0], out_channels, 3, padding=1)

    @classmethod
    def from_config(cls, config=SD15_UNET, **overrides):
        cfg = dict(config)
        cfg.update(overrides)
        return cls(**{k: v for k, v in cfg.items() if not k.startswith("_")})

    @classmethod
    def from_pretrained(cls, path_or_repo: str, subfolder=None, **_unused):
        from .. import hub
        return hub.load_model(cls, path_or_repo, subfolder)

    def save_pretrained(self, directory: str, safe_serialization: bool = True):
        from .. import hub
        hub.save_model(self, directory, "UNet2DConditionModel", safe_serialization)

    @property
    def dtype(self):
        return self.conv_in.weight.dtype

    @property
    def device(self):
        return self.conv_in.weight.device

    # ---- diffusers attention-processor surface (video_equiv_editing_pipeline.py:505-512, :743)
    @property
    def attn_processors(self) -> Dict[str, object]:
        return {f"{name}.processor": m.get_processor() for name, m in self.named_modules()
                if name and hasattr(m, "set_processor") and hasattr(m, "get_processor")}

    def set_attn_processor(self, processor) -> None:
        mods = [(name, m) for name, m in self.named_modules()
                if name and hasattr(m, "set_processor") and hasattr(m, "get_processor")]
        if isinstance(processor, dict):
            if len(processor) != len(mods):
                raise ValueError(f"A dict of processors was passed, but the number of processors {len(processor)} does "
                                 f"not match the number of attention layers: {len(mods)}.")
            for name, m in mods:
                m.set_processor(processor[f"{name}.processor"])
        else:
            for _, m in mods:
                m.set_processor(processor)

    def _resnets(self):
        for blk in self.down_blocks:
            yield from blk.resnets
        yield from self.mid_block.resnets
        for blk in self.up_blocks:
            yield from blk.resnets

    def forward(self, sample: torch.Tensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor = None, return_dict: bool = True, **kwargs):
        bsz = sample.shape[0]
        dev = sample.device
        if torch.is_tensor(timestep):
            t = timestep.to(device=dev, dtype=torch.float32).reshape(-1)
            if t.numel() == 1 and bsz > 1:
                t = t.expand(bsz)
            t = t.contiguous()
        else:
            t = torch.full((bsz,), float(timestep), dtype=torch.float32, device=dev)
        if encoder_hidden_states is None:
            raise ValueError("UNet2DConditionModel needs encoder_hidden_states [B, 77, cross_attention_dim]")
        ehs = encoder_hidden_states.to(device=dev, dtype=torch.float32).contiguous()
        emb = self.time_embedding(ops.timestep_embedding(t, self.config.block_out_channels[0]))
        resnets = list(self._resnets())
        w, b = fused_linear_params(self, "temb_proj", [r.time_emb_proj for r in resnets])
        allp = ops.linear_rows(emb, w, b, act_in="silu")
        projs, off = [], 0
        for r in resnets:
            projs.append(allp[:, off:off + r.out_channels])
            off += r.out_channels

        x = ops.nhwc(sample)
        cin = x.shape[-1]
        if ops.default_conv_algo() == "tf32" and cin % 32 != 0:
            w, b, k = conv_params_padded(self.conv_in, 32 * ((cin + 31) // 32))
            h = ops.nchw_view(ops.conv2d(ops.pad_channels(x, w.shape[-1]), w, b, k, gn_stats=True))
        else:
            w, b, k = conv_params(self.conv_in)
            h = ops.nchw_view(ops.conv2d(x, w, b, k))
        skips = (h,)
        pi = 0
        for blk in self.down_blocks:
            n = len(blk.resnets)
            h, outs = blk(h, emb, projs[pi:pi + n], ehs)
            pi += n
            skips += outs
        h = self.mid_block(h, emb, projs[pi:pi + 2], ehs)
        pi += 2
        for blk in self.up_blocks:
            n = len(blk.resnets)
            h = blk(h, skips[-n:], emb, projs[pi:pi + n], ehs)
            pi += n
            skips = skips[:-n]
        hx = ops.nhwc(h)
        gn = self.conv_norm_out
        a = ops.groupnorm_act(hx, gn.num_groups, gn.eps, gn.weight, gn.bias, act=act_name(self.conv_act))
        w, b, k = conv_params(self.conv_out)
        out = ops.nchw_view(ops.conv2d(a, w, b, k))
        return UNet2DOutput(out) if return_dict else (out,)
