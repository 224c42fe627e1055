"""Oracle (test infrastructure): alias-free blocks and the module-swap surgery, in plain PyTorch.

Follows /root/reference/afldm/af_modules/af_blocks.py:12-152 (WarpedNonlinearity,
AliasFreeUpsample2D, AliasFreeDownsample2D) and /root/reference/afldm/af_modules/af_api.py:9-93
(wrap_nonlinearity, replace_upsampler/downsampler, make_af_unet, make_af_vae[_from_config]).
Operates on the restated diffusers modules of ``oracle.nn``.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ideal_lpf as L


class WarpedNonlinearity(nn.Module):
    """af_blocks.py:12-28: 4-D input -> up2, act, LPF, decimate; lower-rank input -> plain act."""

    def __init__(self, nonlinearity):
        super().__init__()
        self.nonlinearity = nonlinearity

    def forward(self, x):
        if x.ndim < 4:
            return self.nonlinearity(x)
        return L.lpf_rfft(self.nonlinearity(L.upsample_rfft(x, 2)), 0.5)[:, :, ::2, ::2]


class AliasFreeUpsample2D(nn.Module):
    """af_blocks.py:45-106: ideal x2 up-sample, then the original layer's 3x3 conv."""

    def __init__(self, channels, use_conv, out_channels, ori_conv):
        super().__init__()
        self.channels, self.use_conv, self.out_channels = channels, use_conv, out_channels
        self.name, self.norm = "conv", None
        self.conv = ori_conv

    def forward(self, hidden_states, output_size=None, *a, **k):
        assert hidden_states.shape[1] == self.channels
        h = L.upsample_rfft(hidden_states, 2)
        return self.conv(h) if self.use_conv else h


class AliasFreeDownsample2D(nn.Module):
    """af_blocks.py:109-152: conv at stride 1 (:129), LPF, then [::2, ::2]."""

    def __init__(self, channels, use_conv, out_channels, padding, ori_conv):
        super().__init__()
        self.channels, self.use_conv, self.out_channels = channels, use_conv, out_channels
        self.padding, self.name, self.norm = padding, "op", None
        self.conv = ori_conv
        self.conv.stride = (1, 1)

    def forward(self, hidden_states, *a, **k):
        assert hidden_states.shape[1] == self.channels
        h = hidden_states
        if self.use_conv and self.padding == 0:
            h = F.pad(h, (1, 1, 1, 1))
        return L.lpf_rfft(self.conv(h), 0.5)[:, :, ::2, ::2]


# ---------------------------------------------------------------------------- surgery (af_api.py)
def _wrap_block(block):
    for r in block.resnets:
        r.nonlinearity = WarpedNonlinearity(r.nonlinearity)


def _swap_up(block):
    u = block.upsamplers[0]
    block.upsamplers[0] = AliasFreeUpsample2D(u.channels, u.use_conv, u.out_channels, u.conv)


def _swap_down(block):
    d = block.downsamplers[0]
    block.downsamplers[0] = AliasFreeDownsample2D(d.channels, d.use_conv, d.out_channels, d.padding, d.conv)


def make_af_unet(unet):
    """af_api.py:70-83."""
    for b in unet.down_blocks:
        if b.downsamplers is not None:
            _swap_down(b)
        _wrap_block(b)
    _wrap_block(unet.mid_block)
    for b in unet.up_blocks:
        if b.upsamplers is not None:
            _swap_up(b)
        _wrap_block(b)


def make_af_vae(vae, mod_mid_act=True, mod_down_filtered_act=(True,) * 4,
                mod_up_filtered_act=(True,) * 4, mod_resampling_layer=(True,) * 3):
    """af_api.py:34-60 (encoder indexes the resampling flags reversed, :42)."""
    down_flags = list(reversed(list(mod_resampling_layer)))
    for i, b in enumerate(vae.encoder.down_blocks):
        if b.downsamplers is not None and down_flags[i]:
            _swap_down(b)
        if mod_down_filtered_act[i]:
            _wrap_block(b)
    if mod_mid_act:
        _wrap_block(vae.encoder.mid_block)
        _wrap_block(vae.decoder.mid_block)
    for i, b in enumerate(vae.decoder.up_blocks):
        if mod_up_filtered_act[i]:
            _wrap_block(b)
        if b.upsamplers is not None and mod_resampling_layer[i]:
            _swap_up(b)


def make_af_vae_from_config(vae):
    """af_api.py:63-67."""
    c = vae.config
    make_af_vae(vae, c.mid_act, c.down_filtered_act, c.up_filtered_act, c.up_rescale)
