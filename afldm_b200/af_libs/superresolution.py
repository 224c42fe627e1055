"""The x4 degradation operator of the super-resolution scripts, on the sm_100a kernels.

Mirrors the callable surface of /root/reference/afldm/af_libs/superresolution.py: ``build_sr4x(device, sr_filter,
image_size)`` (:288-320, what ``scripts/shift_ldm_sr.py:45,102`` calls), ``build_sr_bicubic`` (:263-280),
``build_sr_pool`` (:283-285) and the ``H / Ht / H_pinv`` methods of the objects they return (:59-86).

The reference keeps the operator as an SVD (``H = U S V^T``, DDRM's ``H_functions``) and multiplies by the three
factors in turn; both of its operators are separable, so here each one is ONE matrix per axis, built on the host in
fp64 and applied per image plane by ``afldm_plane_sep_transform_f32``:

* bicubic (``SRConv``, :160-260): the 1-D matrix H1 [n/4, n] of a 16-tap cubic (a = -0.5) filter at stride 4 with
  reflective padding, its singular values below 3e-2 zeroed (:190-191; none are for the shipped sizes), y = H1 x H1^T;
* pool (``SuperResolution``, :89-157): 4 x 4 mean pooling, H1 = 1-D mean pooling.

``build_sr4x`` follows the degradation with nearest-neighbour x4 up-sampling (:297, :313); the replication matrix is
folded into H1, so ``sr4x(img)`` is a single separable transform.  CUDA tensors only - no CPU fallback.
"""
from __future__ import annotations

import functools

import numpy as np
import torch

from .. import _lib
from .ideal_lpf import sep_transform

_ZERO = 3e-2        # singular values of the 1-D operator below this are dropped (:190)


def _cubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x <= 1.0:
        return (a + 2.0) * x ** 3 - (a + 3.0) * x ** 2 + 1.0
    if x < 2.0:
        return a * x ** 3 - 5.0 * a * x ** 2 + 8.0 * a * x - 4.0 * a
    return 0.0


def bicubic_taps(factor: int) -> np.ndarray:
    """4 * factor taps sampled at ((i - 2 factor) + 1/2) / factor, normalised to sum 1 (:264-276)."""
    k = np.array([_cubic((i - np.floor(factor * 4 / 2) + 0.5) / factor) for i in range(4 * factor)], dtype=np.float64)
    k = (k / k.sum()).astype(np.float32).astype(np.float64)          # the reference normalises, casts to fp32 ...
    return k / k.sum()                                               # ... and normalises again (:277-278)


@functools.lru_cache(maxsize=16)
def conv_matrix(img_dim: int, factor: int) -> np.ndarray:
    """H1 [img_dim / factor, img_dim]: row r filters around input position factor // 2 + r * factor, taps
    j = i - K/2 .. i + K/2 - 1, indices outside the image reflected (-1 -> 0, n -> n - 1) (:175-188), followed by the
    reference's singular-value cut."""
    k = bicubic_taps(factor)
    K = k.shape[0]
    small = img_dim // factor
    h = np.zeros((small, img_dim), dtype=np.float64)
    rows = np.arange(small)
    centres = factor // 2 + rows * factor
    for t in range(K):
        j = centres - K // 2 + t
        j = np.where(j < 0, -j - 1, j)
        j = np.where(j >= img_dim, 2 * img_dim - 1 - j, j)
        np.add.at(h, (rows, j), k[t])
    u, s, vt = np.linalg.svd(h, full_matrices=False)
    if (s < _ZERO).any():
        h = (u * np.where(s < _ZERO, 0.0, s)) @ vt
    return h


@functools.lru_cache(maxsize=16)
def pool_matrix(img_dim: int, factor: int) -> np.ndarray:
    small = img_dim // factor
    return np.kron(np.eye(small), np.full((1, factor), 1.0 / factor))


def _replicate(m: np.ndarray, factor: int) -> np.ndarray:
    return np.repeat(m, factor, axis=0)


class _SeparableH:
    """y = H1 x H1^T per channel, with the reference's flat return shapes (``H_functions.H`` returns [B, C n'^2])."""

    def __init__(self, h1: np.ndarray, channels: int, img_dim: int, ratio: int, device):
        self.h1, self.channels, self.img_dim, self.ratio = h1, channels, img_dim, ratio
        self.small_dim = h1.shape[0]
        self.device = torch.device(device)

    def _planes(self, vec: torch.Tensor, dim: int) -> torch.Tensor:
        if not vec.is_cuda:
            raise _lib.AfldmError("superresolution: CUDA tensors expected (no CPU fallback)")
        return vec.reshape(vec.shape[0], self.channels, dim, dim).to(torch.float32)

    def H(self, vec: torch.Tensor) -> torch.Tensor:
        y = sep_transform(self._planes(vec, self.img_dim), self.h1, self.h1)
        return y.reshape(vec.shape[0], -1)

    def Ht(self, vec: torch.Tensor) -> torch.Tensor:
        y = sep_transform(self._planes(vec, self.small_dim), self.h1.T, self.h1.T)
        return y.reshape(vec.shape[0], -1)

    def H_pinv(self, vec: torch.Tensor) -> torch.Tensor:
        p = np.linalg.pinv(self.h1, rcond=1e-12)
        y = sep_transform(self._planes(vec, self.small_dim), p, p)
        return y.reshape(vec.shape[0], -1)


def build_sr_bicubic(factor, device, image_size, data_channels=3):
    return _SeparableH(conv_matrix(int(image_size), int(factor)), data_channels, int(image_size), int(factor), device)


def build_sr_pool(factor, device, image_size, data_channels=3):
    return _SeparableH(pool_matrix(int(image_size), int(factor)), data_channels, int(image_size), int(factor), device)


def build_sr4x(device, sr_filter, image_size):
    assert sr_filter in ["pool", "bicubic"]
    factor = 4
    h1 = conv_matrix(int(image_size), factor) if sr_filter == "bicubic" else pool_matrix(int(image_size), factor)
    m = _replicate(h1, factor)                      # degradation + nearest x4 back to the input size, one matrix per axis

    def sr4x(img: torch.Tensor) -> torch.Tensor:
        is3d = img.dim() == 3
        if is3d:
            img = img.unsqueeze(0)
        if img.shape[-1] != image_size or img.shape[-2] != image_size:
            raise _lib.AfldmError(f"sr4x: built for {image_size} x {image_size} images, got {tuple(img.shape)}")
        img = img.to(device=device, dtype=torch.float32)
        res = sep_transform(img, m, m)
        return res[0] if is3d else res

    return sr4x
