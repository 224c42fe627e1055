"""Backward of the ideal resamplers and of the filtered activation.

The reference trains through ``torch.fft`` autograd (afldm/trainers/ldm_trainer.py:240-272: the shift-equivariance loss
differentiates ``WarpedNonlinearity`` / ``UpsampleRFFT`` / ``LPF_RFFT``).  The three operators are linear maps per plane,
so their adjoints are the transposed matrices:

    up2(x) = U x U^T                  d/dx:  U^T g U
    lpf_down2(a) = D a D^T            d/da:  D^T g D
    filtered_act(x) = D act(Z) D^T,   Z = U x U^T:   d/dx = U^T [ act'(Z) * (D^T g D) ] U

(``D`` is NOT ``U^T / 2``: the reconstruction mask weights the Nyquist bin 0.5 where the anti-alias mask zeroes it,
ideal_lpf.py:12-49, so each direction uses its own matrix.)  The linear halves run ``afldm_plane_sep_transform_f32`` with
the transposed operators, ``Z`` is recomputed by the fused up-sampler, the elementwise factor is ``afldm_act_bwd_mul_f32``.
These functions make the three STANDALONE modules differentiable (``WarpedNonlinearity``, ``UpsampleRFFT(2)``,
``LPFDown2`` called on a tensor that requires grad); the fused inference path of the models (GroupNorm folded into the
activation, fp16 operand storage, tensor-core convolutions) has no backward - training is outside this build's scope.
"""
from __future__ import annotations

import torch

from .. import _lib, ops
from ..af_libs.ideal_lpf import filter_matrix, sep_transform, upsample_matrix


def _u(n: int):
    return upsample_matrix(n, 2)                       # [2n, n]


def _d(n: int):
    return filter_matrix(2 * n, 0.5, 0.0)[::2, :]      # [n, 2n]


def _nchw(x: torch.Tensor) -> torch.Tensor:
    return ops.to_nchw_contiguous(ops.nhwc(x))


class Up2Ideal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.n = x.shape[-1]
        return ops.nchw_view(ops.up2_ideal(ops.nhwc(x.detach())))

    @staticmethod
    def backward(ctx, g):
        ut = _u(ctx.n).T
        return sep_transform(_nchw(g), ut, ut)


class LpfDown2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        ctx.n = a.shape[-1] // 2
        return ops.nchw_view(ops.lpf_down2(ops.nhwc(a.detach())))

    @staticmethod
    def backward(ctx, g):
        dt = _d(ctx.n).T
        return sep_transform(_nchw(g), dt, dt)


class FilteredAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act: str):
        ctx.save_for_backward(x.detach())
        ctx.act = act
        return ops.nchw_view(ops.filtered_act(ops.nhwc(x.detach()), act=act))

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        n = x.shape[-1]
        z = _nchw(ops.nchw_view(ops.up2_ideal(ops.nhwc(x))))               # Z = U x U^T, recomputed
        dt = _d(n).T
        gz = sep_transform(_nchw(g), dt, dt)                                  # D^T g D on the 2n x 2n plane
        L = _lib.lib()
        out = torch.empty_like(gz)
        _lib.check(L.afldm_act_bwd_mul_f32(z.data_ptr(), gz.data_ptr(), out.data_ptr(), gz.numel(), ops.ACT[ctx.act],
                                           ops._stream()), "act_bwd_mul")
        ut = _u(n).T
        return sep_transform(out, ut, ut), None


def up2_ideal(x: torch.Tensor) -> torch.Tensor:
    return Up2Ideal.apply(x)


def lpf_down2(a: torch.Tensor) -> torch.Tensor:
    return LpfDown2.apply(a)


def filtered_act(x: torch.Tensor, act: str = "silu") -> torch.Tensor:
    return FilteredAct.apply(x, act)
