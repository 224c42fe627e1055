"""Host-side mirror of the diffusers 0.32.1 building blocks the reference mutates, with forwards
that launch the sm_100a kernels (call sites in the reference:
afldm/af_modules/af_api.py:9-31, afldm/pipelines/cross_frame_attn.py:54-130).

Parameter names / shapes are the diffusers ones (``norm1, conv1, time_emb_proj, norm2, conv2,
conv_shortcut``; ``group_norm, to_q, to_k, to_v, to_out.0``; ``downsamplers.0.conv``;
``upsamplers.0.conv``) so a diffusers checkpoint's ``state_dict`` loads unchanged.  ``nn.Conv2d`` /
``nn.GroupNorm`` / ``nn.Linear`` objects are used as parameter containers only; their own
``forward`` is never called on the hot path.

Tensors between modules are logical [B,C,H,W], physically channels_last (NHWC).
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from ..af_libs.ideal_lpf import sep_transform
from ..af_modules.af_blocks import WarpedNonlinearity, act_name
from ..packing import conv_params, conv_params_f16, fused_linear_params, fused_linear_params_f16


def norm_act(x: torch.Tensor, norm: nn.GroupNorm, nonlinearity: nn.Module, out_half: bool = False) -> torch.Tensor:
    """act(GroupNorm(x)) on NHWC x: statistics pass + one fused apply kernel.
    A ``WarpedNonlinearity`` (alias-free surgery) selects the filtered activation.  ``out_half``: the only consumer
    is a tensor-core convolution that accepts fp16 operands - the result MAY then be fp16 (check ``.dtype``)."""
    if isinstance(nonlinearity, WarpedNonlinearity):
        return ops.filtered_act_groupnorm(x, norm.num_groups, norm.eps, norm.weight, norm.bias, act=nonlinearity.act,
                                          out_half=out_half)
    return ops.groupnorm_act(x, norm.num_groups, norm.eps, norm.weight, norm.bias, act=act_name(nonlinearity),
                             out_half=out_half)


def conv_after_act(a: torch.Tensor, conv: nn.Module, **kw) -> torch.Tensor:
    """``ops.conv2d`` of an activation that may have been stored as fp16 (``norm_act(out_half=True)``)."""
    w, b, k = conv_params_f16(conv) if a.dtype == torch.float16 else conv_params(conv)
    return ops.conv2d(a, w, b, k, **kw)


class ResnetBlock2D(nn.Module):
    """diffusers ResnetBlock2D ("default" time-scale-shift, output_scale_factor 1, no dropout)."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int],
                 groups: int = 32, eps: float = 1e-5):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps, affine=True)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, 1, 1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps, affine=True)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, 1, 1)
        self.nonlinearity = nn.SiLU()       # the one shared activation that make_af_* wraps
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, input_tensor: torch.Tensor, temb: Optional[torch.Tensor] = None,
                temb_proj: Optional[torch.Tensor] = None, *args, skip: Optional[torch.Tensor] = None,
                **kwargs) -> torch.Tensor:
        """``temb`` [B, temb_channels]; ``temb_proj`` = precomputed time_emb_proj(act(temb)) [B, Cout]
        (the UNet batches these for all its resnets into one launch).  ``skip``: the block input is
        ``torch.cat([input_tensor, skip], dim=1)`` (up blocks); when both consumers of the concat (norm1 + filtered
        activation, conv_shortcut) can read the two sources directly it is never materialised."""
        x = ops.nhwc(input_tensor)
        if temb_proj is None and temb is not None and self.time_emb_proj is not None:
            temb_proj = ops.linear_rows(temb.contiguous(), self.time_emb_proj.weight, self.time_emb_proj.bias,
                                        act_in="silu")
        bsz, hh, ww = x.shape[0], x.shape[1], x.shape[2]
        cin_total = self.conv1.in_channels
        half1 = ops.conv_f16_supported(bsz, hh, ww, cin_total, self.out_channels)
        half2 = ops.conv_f16_supported(bsz, hh, ww, self.out_channels, self.out_channels)
        act1, sc = None, None
        xs = ops.nhwc(skip) if skip is not None else None
        warped = isinstance(self.nonlinearity, WarpedNonlinearity)
        # conv_shortcut reads the block input only and is consumed by conv2's epilogue: issue it first, on a side stream
        # (a parallel graph branch under norm1 / activation / conv1), with its own split-K scratch buffer
        main = side = None
        if self.conv_shortcut is not None and ops.SHORTCUT_SIDE_STREAM and x.is_cuda and (skip is None or warped):
            main, side = torch.cuda.current_stream(x.device), ops.side_stream(x.device)
            side.wait_stream(main)
            with torch.cuda.stream(side), ops.scratch_slot(ops.side_scratch_slot()):
                ws, bs, ks = conv_params(self.conv_shortcut)
                sc = ops.conv2d_cat(x, xs, ws, bs, ks) if skip is not None else ops.conv2d(x, ws, bs, ks)
        if skip is not None:
            if warped and self.conv_shortcut is not None and (sc is not None or side is None):
                n1 = self.norm1
                act1 = ops.filtered_act_groupnorm_cat(x, xs, n1.num_groups, n1.eps, n1.weight, n1.bias,
                                                      act=self.nonlinearity.act, out_half=half1)
                if act1 is not None and sc is None:
                    ws, bs, ks = conv_params(self.conv_shortcut)
                    sc = ops.conv2d_cat(x, xs, ws, bs, ks)
            if act1 is None or sc is None:
                if side is not None:
                    main.wait_stream(side)          # a shortcut already issued for the un-concatenated input is dropped
                    side = None
                x = ops.concat_channels(x, xs)
                act1, sc = None, None
        if act1 is None:
            act1 = norm_act(x, self.norm1, self.nonlinearity, out_half=half1)
        h = conv_after_act(act1, self.conv1, row_add=temb_proj, gn_stats=True)
        a = norm_act(h, self.norm2, self.nonlinearity, out_half=half2)
        if side is not None:
            main.wait_stream(side)
        if sc is not None:
            out = conv_after_act(a, self.conv2, residual=sc, out=sc, gn_stats=True)
        elif self.conv_shortcut is not None:
            ws, bs, ks = conv_params(self.conv_shortcut)
            sc = ops.conv2d(x, ws, bs, ks)
            out = conv_after_act(a, self.conv2, residual=sc, out=sc, gn_stats=True)
        else:
            out = conv_after_act(a, self.conv2, residual=x, gn_stats=True)
        return ops.nchw_view(out)


class AttnProcessor2_0:
    """diffusers AttnProcessor2_0, CUDA kernels inside.  Two input forms, as in diffusers: the attention-block form
    (4-D [B,C,H,W]: group norm, biased projections, residual) and the transformer form (3-D [B,N,C]: SD-1.5
    ``BasicTransformerBlock.attn1 / attn2``, bias-free q / k / v, no norm, no residual).

    ``encoder_hidden_states`` ([Bkv, Nk, Ckv], already normalised by the caller, as in
    cross_frame_attn.py:86-97) supplies K/V; Bkv may divide B (each K/V batch serves B/Bkv queries
    batches without being tiled in memory)."""

    def __call__(self, attn: "Attention", hidden_states: torch.Tensor,
                 encoder_hidden_states: Optional[torch.Tensor] = None, attention_mask=None, temb=None,
                 *args, **kwargs) -> torch.Tensor:
        if attention_mask is not None:
            raise NotImplementedError("attention_mask is not used on the AF-LDM path")
        spatial = hidden_states.ndim == 4
        if spatial:
            x = ops.nhwc(hidden_states)
        else:
            x3 = hidden_states.contiguous()
            x = x3.view(x3.shape[0], x3.shape[1], 1, x3.shape[2])          # tokens as an [N x 1] image
        b, h, w, c = x.shape
        if attn.to_q.weight.shape[0] != c:
            raise NotImplementedError("attention with inner_dim != query_dim")
        d = c // attn.heads
        # head dims the flash kernels cover: up to 64 in both classes (24 in the LDM UNet, 40 in SD-1.5), 80 / 160 (SD-1.5's
        # deeper levels) on the TF32 tensor-core kernel; the exact-fp32 class runs those as GEMM + row softmax + GEMM
        fast = d % 8 == 0 and (d <= 64 or (d in (80, 160) and ops.default_conv_algo() == "tf32"))
        # TF32 class: the projections may hand q | k | v over as fp16 (same 11-bit significands as TF32 operands)
        # to the ldmatrix / mma.m16n8k16 attention kernel
        f16 = ops.F16_ATTENTION and ops.default_conv_algo() == "tf32" and fast and d <= 64
        # ... and then the normalised input and the attention output, each consumed by ONE projection, are stored as
        # fp16 too (kind::f16 projections: half the operand bytes)
        half = (f16 and encoder_hidden_states is None and attn.group_norm is not None
                and ops.conv_f16_supported(b, h, w, c, 3 * c, 1) and ops.conv_f16_supported(b, h, w, c, c, 1))
        xn = x
        if attn.group_norm is not None:
            gn = attn.group_norm
            xn = ops.groupnorm_act(x, gn.num_groups, gn.eps, gn.weight, gn.bias, act="identity", out_half=half)
        if encoder_hidden_states is None:
            qkv = None
            if xn.dtype == torch.float16:
                wqkv, bqkv = fused_linear_params_f16(attn, "qkv", (attn.to_q, attn.to_k, attn.to_v))
                qkv = ops.conv2d_f16out(xn, wqkv, bqkv, 1)
                if qkv is None:         # split-K shape (small batch): fp32 q | k | v and the TF32-operand attention
                    qkv = ops.conv2d(xn, wqkv, bqkv, 1)
            else:
                wqkv, bqkv = fused_linear_params(attn, "qkv", (attn.to_q, attn.to_k, attn.to_v))
                qkv = ops.conv2d_f16out(xn, wqkv, bqkv, 1) if f16 else None
                if qkv is None:
                    qkv = ops.conv2d(xn, wqkv, bqkv, 1)
            qkv = qkv.view(b, h * w, 3 * c)
            q, k, v = qkv[:, :, :c], qkv[:, :, c:2 * c], qkv[:, :, 2 * c:]
        else:
            src = encoder_hidden_states
            if src.ndim != 3 or src.shape[-1] != attn.to_k.weight.shape[1]:
                raise ValueError("encoder_hidden_states must be [Bkv, Nk, cross_attention_dim]")
            src = src.contiguous()
            wq, bq, _ = conv_params(attn.to_q)
            q = ops.conv2d(xn, wq, bq, 1).view(b, h * w, c)
            wkv, bkv = fused_linear_params(attn, "kv", (attn.to_k, attn.to_v))
            kv = ops.conv2d(src.view(src.shape[0], src.shape[1], 1, src.shape[2]), wkv, bkv, 1)
            kv = kv.view(src.shape[0], src.shape[1], 2 * c)
            k, v = kv[:, :, :c], kv[:, :, c:]
        if q.dtype == torch.float16 and k.dtype == torch.float16:
            o = ops.attention_f16(q, k, v, attn.heads, out_half=xn.dtype == torch.float16)
        elif fast:
            o = ops.attention(q, k, v, attn.heads)
        else:
            o = ops.attention_gemm(q, k, v, attn.heads)
        out = conv_after_act(o.view(b, h, w, c), attn.to_out[0], residual=x if attn.residual_connection else None,
                             gn_stats=spatial)
        if attn.rescale_output_factor != 1.0:
            raise NotImplementedError("rescale_output_factor != 1")
        return ops.nchw_view(out) if spatial else out.view(b, h * w, c)


class Attention(nn.Module):
    """diffusers Attention: the attention-block configuration (bias, group norm, residual; the default) or the
    transformer configuration of SD-1.5 (``norm=False, residual=False, bias=False``, optional ``cross_attention_dim``)."""

    def __init__(self, channels: int, heads: int, dim_head: int, groups: int = 32, eps: float = 1e-5,
                 cross_attention_dim: Optional[int] = None, bias: bool = True, norm: bool = True, residual: bool = True):
        super().__init__()
        if heads * dim_head != channels:
            raise ValueError("heads * dim_head must equal channels")
        self.heads, self.dim_head = heads, dim_head
        self.scale = dim_head ** -0.5
        self.residual_connection = residual
        self.rescale_output_factor = 1.0
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True) if norm else None
        kv_dim = cross_attention_dim or channels
        self.to_q = nn.Linear(channels, channels, bias=bias)
        self.to_k = nn.Linear(kv_dim, channels, bias=bias)
        self.to_v = nn.Linear(kv_dim, channels, bias=bias)
        self.to_out = nn.ModuleList([nn.Linear(channels, channels), nn.Dropout(0.0)])
        self.processor = AttnProcessor2_0()

    def set_processor(self, processor) -> None:
        self.processor = processor

    def get_processor(self):
        return self.processor

    def forward(self, hidden_states, encoder_hidden_states=None, **kwargs):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, **kwargs)


def _decimate2(x_nchw: torch.Tensor, phase: int) -> torch.Tensor:
    """x[:, :, phase::2, phase::2] as a selection operator on ``afldm_plane_sep_transform_f32``."""
    h, w = x_nchw.shape[-2:]
    sel_h = np.eye(h)[phase::2][: h // 2]
    sel_w = np.eye(w)[phase::2][: w // 2]
    return sep_transform(x_nchw, sel_h, sel_w)


class Downsample2D(nn.Module):
    """diffusers Downsample2D (3x3 conv, stride 2).  Holds the weights that
    ``replace_downsampler`` hands to ``AliasFreeDownsample2D``."""

    def __init__(self, channels: int, use_conv: bool = True, out_channels: Optional[int] = None,
                 padding: int = 1, name: str = "op"):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.padding, self.name = use_conv, padding, name
        self.norm = None
        self.conv = nn.Conv2d(channels, self.out_channels, 3, stride=2, padding=padding)

    def forward(self, hidden_states, *args, **kwargs):
        """The aliasing original (3x3 conv, stride 2): the stride-1 'same' convolution sampled on the stride-2 grid -
        even positions for padding 1, odd positions for the padding-0 form that pads (0, 1, 0, 1) first (diffusers
        Downsample2D).  Not on the alias-free path (surgery replaces it); kept so that a model can be
        run before ``make_af_*`` for an A/B of the aliasing baseline."""
        w, b, k = conv_params(self.conv)
        h = ops.conv2d(ops.nhwc(hidden_states), w, b, k)
        return _decimate2(ops.to_nchw_contiguous(h), 0 if self.padding == 1 else 1)


class Upsample2D(nn.Module):
    """diffusers Upsample2D (nearest x2 + 3x3 conv); see ``Downsample2D``."""

    def __init__(self, channels: int, use_conv: bool = True, out_channels: Optional[int] = None,
                 name: str = "conv"):
        super().__init__()
        self.channels, self.out_channels = channels, out_channels or channels
        self.use_conv, self.name = use_conv, name
        self.use_conv_transpose = False
        self.norm = None
        self.conv = nn.Conv2d(channels, self.out_channels, 3, padding=1)

    def forward(self, hidden_states, output_size=None, *args, **kwargs):
        """The aliasing original: nearest-neighbour x2 (``F.interpolate`` in diffusers) followed by the 3x3 conv."""
        x = ops.to_nchw_contiguous(ops.nhwc(hidden_states))
        n = x.shape[-1]
        rep = np.repeat(np.eye(n), 2, axis=0)                 # [2n, n]: row o copies sample o // 2
        up = sep_transform(x, np.repeat(np.eye(x.shape[-2]), 2, axis=0), rep)
        if not self.use_conv:
            return up
        w, b, k = conv_params(self.conv)
        return ops.nchw_view(ops.conv2d(ops.nhwc(up), w, b, k, gn_stats=True))


# --------------------------------------------------------------------------------- UNet blocks
class DownBlock2D(nn.Module):
    """DownBlock2D / AttnDownBlock2D (``attention_head_dim`` None = no attention)."""

    def __init__(self, in_channels, out_channels, temb_channels, num_layers, add_downsample,
                 attention_head_dim, eps, downsample_padding=1):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps=eps)
            for i in range(num_layers)])
        self.attentions = None
        if attention_head_dim:
            self.attentions = nn.ModuleList([
                Attention(out_channels, out_channels // attention_head_dim, attention_head_dim, eps=eps)
                for _ in range(num_layers)])
        self.downsamplers = None
        if add_downsample:
            self.downsamplers = nn.ModuleList(
                [Downsample2D(out_channels, True, out_channels, padding=downsample_padding, name="op")])

    def forward(self, hidden_states, temb=None, temb_projs=None):
        outputs = ()
        for i, resnet in enumerate(self.resnets):
            hidden_states = resnet(hidden_states, temb, None if temb_projs is None else temb_projs[i])
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states)
            outputs += (hidden_states,)
        if self.downsamplers is not None:
            hidden_states = self.downsamplers[0](hidden_states)
            outputs += (hidden_states,)
        return hidden_states, outputs


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels, temb_channels, attention_head_dim, eps):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, temb_channels, eps=eps),
                                      ResnetBlock2D(channels, channels, temb_channels, eps=eps)])
        self.attentions = nn.ModuleList(
            [Attention(channels, channels // attention_head_dim, attention_head_dim, eps=eps)])

    def forward(self, hidden_states, temb=None, temb_projs=None):
        tp = temb_projs if temb_projs is not None else (None, None)
        hidden_states = self.resnets[0](hidden_states, temb, tp[0])
        hidden_states = self.attentions[0](hidden_states)
        return self.resnets[1](hidden_states, temb, tp[1])


class UpBlock2D(nn.Module):
    """UpBlock2D / AttnUpBlock2D: each resnet consumes cat([h, skip]) (skip popped from the end)."""

    def __init__(self, in_channels, out_channels, prev_output_channel, temb_channels, num_layers,
                 add_upsample, attention_head_dim, eps):
        super().__init__()
        resnets = []
        for i in range(num_layers):
            skip_channels = in_channels if i == num_layers - 1 else out_channels
            first = prev_output_channel if i == 0 else out_channels
            resnets.append(ResnetBlock2D(first + skip_channels, out_channels, temb_channels, eps=eps))
        self.resnets = nn.ModuleList(resnets)
        self.attentions = None
        if attention_head_dim:
            self.attentions = nn.ModuleList([
                Attention(out_channels, out_channels // attention_head_dim, attention_head_dim, eps=eps)
                for _ in range(num_layers)])
        self.upsamplers = None
        if add_upsample:
            self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)])

    def forward(self, hidden_states, res_hidden_states_tuple, temb=None, temb_projs=None):
        skips = list(res_hidden_states_tuple)
        for i, resnet in enumerate(self.resnets):
            skip = skips.pop()
            hidden_states = resnet(hidden_states, temb, None if temb_projs is None else temb_projs[i], skip=skip)
            if self.attentions is not None:
                hidden_states = self.attentions[i](hidden_states)
        if self.upsamplers is not None:
            hidden_states = self.upsamplers[0](hidden_states)
        return hidden_states


# --------------------------------------------------------------------------------- VAE blocks
class UpDecoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, None, eps=eps)
            for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels, True, out_channels)]) if add_upsample else None

    def forward(self, hidden_states, temb=None):
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, None)
        if self.upsamplers is not None:
            hidden_states = self.upsamplers[0](hidden_states)
        return hidden_states


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample, eps=1e-6):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, None, eps=eps)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList(
            [Downsample2D(out_channels, True, out_channels, padding=0, name="op")]) if add_downsample else None

    def forward(self, hidden_states):
        for resnet in self.resnets:
            hidden_states = resnet(hidden_states, None)
        if self.downsamplers is not None:
            hidden_states = self.downsamplers[0](hidden_states)
        return hidden_states
