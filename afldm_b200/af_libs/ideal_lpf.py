"""Module mirror of the reference's ideal-filter layers, running the sm_100a kernels.

Same names, constructor arguments and call signatures as /root/reference/afldm/af_libs/ideal_lpf.py
(``create_lpf_rect`` :12-24, ``create_recon_rect`` :38-49, ``LPF_RFFT`` :52-93, ``LPF_RECON_RFFT`` :96-134,
``UpsampleRFFT`` :137-158, ``subpixel_shift`` :161-172) on logical [B,C,H,W] fp32 CUDA tensors.

Where the reference runs ``rfft2 -> mask -> irfft2``, the same linear operator is applied here in its
separable matrix form ``y = F x F^T`` per plane (SURVEY.md 8(a) identity 2; F is the inverse DFT of the 1-D
mask as a circulant, built in fp64 and cached per (size, cutoff)):

* the two configurations of the hot path - ``UpsampleRFFT(up=2)`` and ``LPF_RFFT(0.5)`` followed by
  ``[::2, ::2]`` (``LPFDown2``) - run the fused resampling kernels (``afldm_up2_ideal_f32`` /
  ``afldm_lpf_down2_f32``: power-of-two square planes up to 128);
* every other configuration (any cutoff, any ``up``, any square plane size, ``subpixel_shift``) runs
  ``afldm_plane_sep_transform_f32`` with the operator matrices - the r-fold zero-stuffed tensor of
  ``UpsampleRFFT.forward`` (:150-155) is never materialised.

Like the reference, the masks are built from the LAST dimension only (:81), so planes must be square.
There is no CPU / cuFFT fallback.
"""
from __future__ import annotations

import functools

import numpy as np
import torch
import torch.nn as nn

from .. import _lib, ops


def _band_1d(n: int, cutoff: float, edge: float) -> np.ndarray:
    """1-D mask over DFT bins 0..n-1 (ideal_lpf.py:12-24 / :38-49): pass ``k <= c`` and ``k >= n - c`` with
    ``c = floor(n * cutoff / 2)``; when ``n % 4 == 0`` the two edge bins get ``edge`` (0 = anti-alias, 0.5 = reconstruction)."""
    c = int((n * cutoff) // 2)
    m = np.ones(n, dtype=np.float64)
    m[c + 1:n - c] = 0.0
    if n % 4 == 0:
        m[c] = edge
        m[n - c] = edge
    return m


def create_lpf_rect(N: int, cutoff: float = 0.5) -> torch.Tensor:
    m = torch.from_numpy(_band_1d(N, cutoff, 0.0)).to(torch.float32)
    return m[:, None] * m[None, :]


def create_recon_rect(N: int, cutoff: float = 0.5) -> torch.Tensor:
    m = torch.from_numpy(_band_1d(N, cutoff, 0.5)).to(torch.float32)
    return m[:, None] * m[None, :]


@functools.lru_cache(maxsize=64)
def filter_matrix(n: int, cutoff: float, edge: float) -> np.ndarray:
    """F in R^{n x n} (fp64): ``irfft(rfft(x) * mask)`` of a length-n signal as a circulant, F[i, j] = f[(i - j) mod n]."""
    f = np.real(np.fft.ifft(_band_1d(n, cutoff, edge)))
    idx = (np.arange(n)[:, None] - np.arange(n)[None, :]) % n
    return f[idx]


@functools.lru_cache(maxsize=64)
def upsample_matrix(n: int, up: int, factor: float = 1.0) -> np.ndarray:
    """U in R^{n up x n}: zero insertion (:150-155), reconstruction filter at cutoff factor / up, gain ``up`` per axis."""
    return filter_matrix(n * up, 1.0 / up * factor, 0.5)[:, ::up] * float(up)


def sep_transform(x: torch.Tensor, my: np.ndarray, mx: np.ndarray) -> torch.Tensor:
    """y = My x Mx^T per plane of the NCHW tensor x (one launch pair, ``afldm_plane_sep_transform_f32``)."""
    if not x.is_cuda or x.dtype != torch.float32 or x.ndim != 4:
        raise _lib.AfldmError("ideal_lpf: fp32 CUDA [B,C,H,W] tensor expected (no CPU fallback)")
    x = x.contiguous()
    b, c, h, w = x.shape
    ho, wo = my.shape[0], mx.shape[0]
    my_d = torch.from_numpy(np.ascontiguousarray(my, dtype=np.float32)).to(x.device).view(1, ho, h)
    mx_d = torch.from_numpy(np.ascontiguousarray(mx, dtype=np.float32)).to(x.device).view(1, wo, w)
    out = torch.empty((b, c, ho, wo), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    planes = b * c
    need = L.afldm_plane_sep_transform_workspace_floats(planes, h, wo)
    ws = ops.scratch(x.device, need)
    ops._run("plane_sep_transform", dict(elems=planes * ho * wo),
             lambda: L.afldm_plane_sep_transform_f32(x.data_ptr(), my_d.data_ptr(), mx_d.data_ptr(), out.data_ptr(),
                                                     ws.data_ptr(), need, planes, planes, h, w, ho, wo, ops._stream()),
             (x, my_d, mx_d, out, ws))
    return out


def _square(x: torch.Tensor, who: str) -> int:
    if x.ndim != 4 or x.shape[-1] != x.shape[-2]:
        raise _lib.AfldmError(f"{who}: square [B,C,N,N] planes expected (the reference builds its mask from the last "
                              f"dimension only, ideal_lpf.py:81), got {tuple(x.shape)}")
    return x.shape[-1]


def _fused_ok(n: int, c: int) -> bool:
    """The fused resampling kernels take power-of-two planes up to 128 and channel counts that are multiples of 32."""
    return n in (2, 4, 8, 16, 32, 64, 128) and c % 32 == 0


class LPF_RFFT(nn.Module):
    """ideal_lpf.py:52-93: full-resolution ideal low-pass at ``cutoff``."""

    def __init__(self, cutoff=0.5, transform_mode="rfft", fixed_size=None):
        super().__init__()
        if transform_mode not in ("fft", "rfft"):
            raise AssertionError(f"transform_mode={transform_mode} is not supported")
        self.cutoff, self.transform_mode, self.fixed_size = cutoff, transform_mode, fixed_size

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = _square(x, "LPF_RFFT")
        f = filter_matrix(n, float(self.cutoff), 0.0)
        return sep_transform(x, f, f)


class LPF_RECON_RFFT(nn.Module):
    """ideal_lpf.py:96-134: reconstruction low-pass (edge bins weighted 0.5)."""

    def __init__(self, cutoff=0.5, transform_mode="rfft"):
        super().__init__()
        if transform_mode not in ("fft", "rfft"):
            raise AssertionError(f"mode={transform_mode} is not supported")
        self.cutoff, self.transform_mode = cutoff, transform_mode

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = _square(x, "LPF_RECON_RFFT")
        f = filter_matrix(n, float(self.cutoff), 0.5)
        return sep_transform(x, f, f)


class UpsampleRFFT(nn.Module):
    """ideal_lpf.py:137-158: y = U x U^T per plane (exact circular sinc interpolation by ``up``)."""

    def __init__(self, up=2, transform_mode="rfft", factor=1):
        super().__init__()
        self.up, self.factor = int(up), float(factor)
        self.recon_filter = LPF_RECON_RFFT(cutoff=1 / up * factor, transform_mode=transform_mode)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n = _square(x, "UpsampleRFFT")
        if self.up == 2 and self.factor == 1.0 and _fused_ok(n, x.shape[1]):
            if torch.is_grad_enabled() and x.requires_grad:
                from ..af_modules.autograd import up2_ideal
                return up2_ideal(x)
            return ops.nchw_view(ops.up2_ideal(ops.nhwc(x)))
        if self.up == 1 and self.factor == 1.0:
            return self.recon_filter(x)
        u = upsample_matrix(n, self.up, self.factor)
        return sep_transform(x, u, u)


class LPFDown2(nn.Module):
    """``LPF_RFFT(0.5)(x)[:, :, ::2, ::2]`` (af_blocks.py:26-27, 149-150) as one fused op."""

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        n2 = _square(x, "LPFDown2")
        if n2 % 2 == 0 and _fused_ok(n2 // 2, x.shape[1]):
            if torch.is_grad_enabled() and x.requires_grad:
                from ..af_modules.autograd import lpf_down2
                return lpf_down2(x)
            return ops.nchw_view(ops.lpf_down2(ops.nhwc(x)))
        d = filter_matrix(n2, 0.5, 0.0)[::2, :]
        return sep_transform(x, d, d)


def subpixel_shift(images: torch.Tensor, up=2, shift_x=1, shift_y=1, up_method="ideal") -> torch.Tensor:
    """ideal_lpf.py:161-172: effective fractional shift (shift_x / up, shift_y / up): up-sample, roll by
    (-shift_x, -shift_y), decimate - folded into one operator matrix per axis."""
    assert up_method == "ideal", 'Only "ideal" interpolation kenrel is supported'
    n = _square(images, "subpixel_shift")
    u = upsample_matrix(n, int(up)) if int(up) > 1 else np.eye(n)
    m = n * int(up)
    my = u[(np.arange(n) * int(up) + int(shift_x)) % m, :]
    mx = u[(np.arange(n) * int(up) + int(shift_y)) % m, :]
    return sep_transform(images, my, mx)
