"""Module mirror of the reference's ideal-filter layers, running the sm_100a kernels.

Same class names and call signatures as /root/reference/afldm/af_libs/ideal_lpf.py:52-158
(``LPF_RFFT(cutoff)(x)``, ``UpsampleRFFT(up)(x)`` on logical [B,C,H,W] tensors); only the
configurations the alias-free blocks use are provided natively: ``UpsampleRFFT(up=2)`` and
``LPF_RFFT(cutoff=0.5)`` followed by the ``[::2, ::2]`` decimation (``lpf_down2``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .. import ops


class UpsampleRFFT(nn.Module):
    """ideal_lpf.py:137-158 for ``up == 2``: y = U x U^T per plane (exact circular sinc)."""

    def __init__(self, up: int = 2, factor: float = 1.0):
        super().__init__()
        if up != 2 or factor != 1.0:
            raise NotImplementedError("afldm_b200 provides the x2 ideal up-sampler (up=2, factor=1)")
        self.up = up

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.nchw_view(ops.up2_ideal(ops.nhwc(x)))


class LPFDown2(nn.Module):
    """``LPF_RFFT(0.5)(x)[:, :, ::2, ::2]`` (af_blocks.py:26-27, 149-150) as one fused op."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.nchw_view(ops.lpf_down2(ops.nhwc(x)))
