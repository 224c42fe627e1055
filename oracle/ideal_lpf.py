"""Oracle (test infrastructure): ideal circular low-pass / re-sampling operators.

Two independent restatements of /root/reference/afldm/af_libs/ideal_lpf.py:

1. FFT form (torch fp32) - the same algorithm the reference runs:
   rfft2 -> multiply by a 0 / 0.5 / 1 rectangle -> irfft2 (ideal_lpf.py:69-93, 112-134, 148-158).
2. Separable operator form (numpy fp64): ``up2(x) = U x U^T`` and ``lpf_down2(y) = D y D^T``
   with closed-form circulant matrices (SURVEY.md section 8(a) identities 2 and 4).  This is
   the formulation the CUDA kernels implement, so it is pinned against (1) in the tests.
"""
from __future__ import annotations

import functools

import numpy as np
import torch


# --------------------------------------------------------------------------- masks (a1, a2)
def _band_1d(n: int, cutoff: float, nyquist_value: float) -> torch.Tensor:
    """1-D frequency mask over fft bins 0..n-1 (ideal_lpf.py:12-24 / 38-49).

    Bins ``k <= c`` and ``k >= n - c`` pass (``c = floor(n * cutoff / 2)``); when ``n % 4 == 0``
    the two edge bins ``c`` and ``n - c`` are set to ``nyquist_value`` (0 for the anti-alias
    filter, 0.5 for the reconstruction filter).
    """
    c = int((n * cutoff) // 2)
    k = torch.arange(n)
    keep = ((k <= c) | (k >= n - c)).to(torch.float32)
    if n % 4 == 0:
        keep[c] = nyquist_value
        keep[n - c] = nyquist_value
    return keep


def lpf_rect(n: int, cutoff: float = 0.5) -> torch.Tensor:
    """`create_lpf_rect` (ideal_lpf.py:12-24): 2-D outer product, Nyquist edge zeroed."""
    m = _band_1d(n, cutoff, 0.0)
    return m[:, None] * m[None, :]


def recon_rect(n: int, cutoff: float = 0.5) -> torch.Tensor:
    """`create_recon_rect` (ideal_lpf.py:38-49): Nyquist edge weighted 0.5."""
    m = _band_1d(n, cutoff, 0.5)
    return m[:, None] * m[None, :]


# --------------------------------------------------------------------------- FFT form (a3-a5)
def lpf_rfft(x: torch.Tensor, cutoff: float = 0.5) -> torch.Tensor:
    """`LPF_RFFT.forward` (ideal_lpf.py:69-93).  Mask size follows the LAST dim (``N = W``)."""
    h, w = x.shape[-2], x.shape[-1]
    mask = lpf_rect(w, cutoff)[:, : w // 2 + 1].to(x.device)
    return torch.fft.irfft2(torch.fft.rfft2(x) * mask, s=(h, w))


def lpf_recon_rfft(x: torch.Tensor, cutoff: float = 0.5) -> torch.Tensor:
    """`LPF_RECON_RFFT.forward` (ideal_lpf.py:112-134); irfft2 with default (even) size."""
    w = x.shape[-1]
    mask = recon_rect(w, cutoff)[:, : w // 2 + 1].to(x.device)
    return torch.fft.irfft2(torch.fft.rfft2(x) * mask)


def zero_insert(x: torch.Tensor, up: int) -> torch.Tensor:
    """Zero-stuffing of ideal_lpf.py:150-155: sample (i, j) lands on (i*up, j*up)."""
    b, c, h, w = x.shape
    y = x.new_zeros(b, c, h * up, w * up)
    y[:, :, ::up, ::up] = x
    return y


def upsample_rfft(x: torch.Tensor, up: int = 2, factor: float = 1.0) -> torch.Tensor:
    """`UpsampleRFFT.forward` (ideal_lpf.py:148-158)."""
    return lpf_recon_rfft(zero_insert(x, up), cutoff=1.0 / up * factor) * float(up * up)


def subpixel_shift(x: torch.Tensor, up: int = 2, shift_x: int = 1, shift_y: int = 1) -> torch.Tensor:
    """`subpixel_shift` (ideal_lpf.py:161-172)."""
    big = upsample_rfft(x, up)
    return torch.roll(big, shifts=(-shift_x, -shift_y), dims=(2, 3))[:, :, ::up, ::up]


def filtered_act_fft(x: torch.Tensor, act=torch.nn.functional.silu) -> torch.Tensor:
    """`WarpedNonlinearity.forward` for 4-D input (af_blocks.py:19-28)."""
    return lpf_rfft(act(upsample_rfft(x, 2)), 0.5)[:, :, ::2, ::2]


# --------------------------------------------------------------------------- operator form
@functools.lru_cache(maxsize=None)
def up2_matrix(n: int) -> np.ndarray:
    """U in R^{2n x n}: the 1-D action of `UpsampleRFFT(2)` on a length-n circular signal.

    Built from the definition (zero-insert, DFT, recon mask, inverse DFT, x2 per axis) in fp64;
    even rows are the identity, odd rows a circulant (SURVEY.md 8(a) identity 2).
    """
    m = 2 * n
    mask = _band_1d(m, 0.5, 0.5).double().numpy()
    eye = np.zeros((m, n))
    eye[::2, :] = np.eye(n)
    spec = np.fft.fft(eye, axis=0) * mask[:, None]
    return np.ascontiguousarray(np.real(np.fft.ifft(spec, axis=0)) * 2.0)


@functools.lru_cache(maxsize=None)
def down2_matrix(n: int) -> np.ndarray:
    """D in R^{n x 2n}: 1-D action of `LPF_RFFT(0.5)` on length 2n followed by ``[::2]``."""
    m = 2 * n
    mask = _band_1d(m, 0.5, 0.0).double().numpy()
    spec = np.fft.fft(np.eye(m), axis=0) * mask[:, None]
    full = np.real(np.fft.ifft(spec, axis=0))
    return np.ascontiguousarray(full[::2, :])


@functools.lru_cache(maxsize=None)
def lpf_matrix(n: int) -> np.ndarray:
    """L in R^{n x n}: 1-D action of `LPF_RFFT(0.5)` (no decimation)."""
    mask = _band_1d(n, 0.5, 0.0).double().numpy()
    spec = np.fft.fft(np.eye(n), axis=0) * mask[:, None]
    return np.ascontiguousarray(np.real(np.fft.ifft(spec, axis=0)))


def up2_taps(n: int) -> np.ndarray:
    """Closed form of the odd-phase circulant: d[r] = (1/n)(1 + 2 sum_{k=1}^{n/2-1} cos(2 pi k (r+1/2)/n))
    plus the half-weighted Nyquist term; returned as the first odd row of ``up2_matrix`` re-indexed
    so that ``U[2i+1, j] = d[(i - j) mod n]``."""
    u = up2_matrix(n)
    return np.array([u[1, (-r) % n] for r in range(n)])


def _apply_sep(x: np.ndarray, mh: np.ndarray, mw: np.ndarray) -> np.ndarray:
    # x: (..., H, W) ; out[..., i, j] = sum_{h,w} mh[i,h] x[...,h,w] mw[j,w]
    return np.einsum("ih,...hw,jw->...ij", mh, x, mw, optimize=True)


def up2_sep(x: np.ndarray) -> np.ndarray:
    h, w = x.shape[-2:]
    return _apply_sep(x.astype(np.float64), up2_matrix(h), up2_matrix(w))


def lpf_down2_sep(y: np.ndarray) -> np.ndarray:
    h, w = y.shape[-2:]
    return _apply_sep(y.astype(np.float64), down2_matrix(h // 2), down2_matrix(w // 2))


def filtered_act_sep(x: np.ndarray) -> np.ndarray:
    """D silu(U x U^T) D^T per plane, fp64 (SURVEY.md 8(a) identity 4)."""
    z = up2_sep(x)
    a = z / (1.0 + np.exp(-z))
    return lpf_down2_sep(a)
