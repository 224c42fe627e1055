"""Alias-free blocks with the reference's module surface, backed by the sm_100a kernels.

Mirrors /root/reference/afldm/af_modules/af_blocks.py:
  * ``WarpedNonlinearity``      (:12-28)   up2 -> act -> LPF -> [::2, ::2], one fused kernel
  * ``AliasFreeUpsample2D``     (:45-106)  ideal x2 up-sample -> the original layer's 3x3 conv
  * ``AliasFreeDownsample2D``   (:109-152) conv at stride 1 -> LPF -> [::2, ::2]
Inputs / outputs are logical [B,C,H,W] fp32 CUDA tensors (any memory format in, channels_last
out).  The conv weights stay ``nn.Parameter``s of the original ``conv`` module, so ``state_dict()``
keys are unchanged (``...samplers.0.conv.{weight,bias}``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops
from ..packing import conv_params, conv_params_f16


def act_name(nonlinearity: nn.Module) -> str:
    if isinstance(nonlinearity, nn.SiLU):
        return "silu"
    if isinstance(nonlinearity, nn.Identity):
        return "identity"
    raise NotImplementedError(f"no CUDA kernel for activation {type(nonlinearity).__name__} (act_fn: silu expected)")


class WarpedNonlinearity(nn.Module):
    def __init__(self, nonlinearity: nn.Module):
        super().__init__()
        self.nonlinearity = nonlinearity
        self.act = act_name(nonlinearity)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.ndim < 4:                      # the time-embedding vector: plain activation (:20-21)
            flat = x.reshape(1, 1, -1)
            if flat.shape[-1] % 4 == 0 and x.is_contiguous():
                return ops.affine_act(flat, None, None, self.act).view(x.shape)
            return self.nonlinearity(x)
        if torch.is_grad_enabled() and x.requires_grad:            # differentiable form
            from .autograd import filtered_act
            return filtered_act(x, self.act)
        return ops.nchw_view(ops.filtered_act(ops.nhwc(x), act=self.act))


class AliasFreeUpsample2D(nn.Module):
    def __init__(self, channels: int, use_conv: bool = False, use_conv_transpose: bool = False,
                 out_channels: int = None, name: str = "conv", ori_conv: nn.Module = None, **_unused):
        super().__init__()
        if use_conv_transpose:
            raise NotImplementedError("use_conv_transpose is not on the AF-LDM path")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_conv_transpose = False
        self.name = name
        self.norm = None
        self.interpolate = True
        self.conv = ori_conv

    def forward(self, hidden_states: torch.Tensor, output_size=None, *args, **kwargs) -> torch.Tensor:
        assert hidden_states.shape[1] == self.channels
        x = ops.nhwc(hidden_states)
        if not self.use_conv:
            return ops.nchw_view(ops.up2_ideal(x))
        # TF32 class: the up-sampled tensor is read by the convolution only - stored as fp16 (kind::f16 operands)
        half = ops.conv_f16_supported(x.shape[0], 2 * x.shape[1], 2 * x.shape[2], self.channels, self.out_channels)
        h = ops.up2_ideal(x, out_half=half)
        w, b, k = conv_params_f16(self.conv) if h.dtype == torch.float16 else conv_params(self.conv)
        h = ops.conv2d(h, w, b, k, gn_stats=True)           # feeds the next block's GroupNorm (via the skip concat)
        return ops.nchw_view(h)


class AliasFreeDownsample2D(nn.Module):
    def __init__(self, channels: int, use_conv: bool = False, out_channels: int = None, padding: int = 1,
                 name: str = "conv", ori_conv: nn.Module = None, **_unused):
        super().__init__()
        if not use_conv:
            raise NotImplementedError("the alias-free down-sampler is conv -> LPF -> decimate (use_conv=True)")
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.padding = padding
        self.name = name
        self.norm = None
        self.conv = ori_conv
        self.conv.stride = (1, 1)           # af_blocks.py:129

    def forward(self, hidden_states: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        assert hidden_states.shape[1] == self.channels
        # padding == 0 (VAE encoder): the reference zero-pads (1,1,1,1) and runs the pad-0 conv (:142-144);
        # padding == 1: the conv pads itself.  Both are the same "same"-size 3x3 convolution.
        w, b, k = conv_params(self.conv)
        h = ops.conv2d(ops.nhwc(hidden_states), w, b, k)
        return ops.nchw_view(ops.lpf_down2(h, gn_stats=True))
