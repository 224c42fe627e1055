"""Fractional-shift warps of the equivariance measurement, on device.

Mirror of ``ImageShifter`` for the ideal filters (/root/reference/afldm/shift_utils/shifters.py:141-191:
``ImageShifter('ideal' | 'ideal_crop', upsample_ratio).shift(img, ti, tj) -> (warped, mask)``) and of
``gen_valid_mask`` (:31-49).  The reference up-samples the image r-fold with ``UpsampleRFFT(r)`` (64x the data at
r = 8), rolls it by ``(round(ti r), round(tj r))``, masks and decimates.  Every step is linear and separable, so
per axis the chain is one n x n matrix

    S[i, j] = U_r[(i r - s) mod (n r), j]          (row i zeroed where the up-sampled validity mask is 0)

with ``U_r`` the 1-D action of ``UpsampleRFFT(r)`` (zero insertion, reconstruction mask of
ideal_lpf.py:38-49 with cutoff 1/r, inverse DFT, gain r), built here in fp64.  The warp is then
``y = S_i x S_j^T`` per plane - ``afldm_plane_sep_transform_f32`` - and a whole sweep of shifts
(``shift_batch``; scripts/shift_ldm_ffhq.py loops over 16 of them) is one launch pair.

The metric code (``mask_mse`` / ``mask_psnr``, metrics.py:5-15) stays in PyTorch, as SURVEY.md 8(a) a15 asks.
"""
from __future__ import annotations

import functools
from typing import Sequence, Tuple

import numpy as np
import torch

from .. import _lib, ops


def gen_valid_mask(shape, ti, tj) -> torch.Tensor:
    """shifters.py:31-49: ones, with the rows / columns the shift wrapped around set to 0."""
    _, _, h, w = shape
    i1, i2 = (0, int(np.ceil(ti))) if ti >= 0 else (int(np.floor(ti)), h)
    j1, j2 = (0, int(np.ceil(tj))) if tj >= 0 else (int(np.floor(tj)), w)
    mask = torch.ones(shape, dtype=torch.float32)
    mask[:, :, i1:i2, :] = 0
    mask[:, :, :, j1:j2] = 0
    return mask


@functools.lru_cache(maxsize=32)
def _upsample_matrix(n: int, r: int) -> np.ndarray:
    """U_r in R^{nr x n} (fp64): 1-D action of UpsampleRFFT(r) on a circular length-n signal
    (ideal_lpf.py:148-158 with the reconstruction mask of :38-49 at cutoff 1/r)."""
    m = n * r
    c = int((m * (1.0 / r)) // 2)
    k = np.arange(m)
    mask = ((k <= c) | (k >= m - c)).astype(np.float64)
    if m % 4 == 0:
        mask[c] = 0.5
        mask[m - c] = 0.5
    eye = np.zeros((m, n))
    eye[::r, :] = np.eye(n)
    spec = np.fft.fft(eye, axis=0) * mask[:, None]
    return np.real(np.fft.ifft(spec, axis=0)) * float(r)


def shift_matrix(n: int, t: float, r: int, crop: bool) -> np.ndarray:
    """The n x n fp64 matrix of one axis of ImageShifter(...).shift for a shift of t pixels at ratio r."""
    s = int(np.round(t * r))
    if r == 1:
        u = np.eye(n)
    else:
        u = _upsample_matrix(n, r)
    m = n * r
    rows = (np.arange(n) * r - s) % m
    out = u[rows, :].copy()
    if crop:
        valid = np.ones(m)
        if s >= 0:
            valid[0:s] = 0            # gen_valid_mask on the up-sampled grid: [0, ceil(s))
        else:
            valid[s:] = 0             # [floor(s), m)
        out *= valid[np.arange(n) * r][:, None]
    return out


class ImageShifter:
    """``ImageShifter(filter, upsample_ratio)`` for filter in {'ideal', 'ideal_crop'} (shifters.py:141-156)."""

    def __init__(self, filter: str = "ideal_crop", upsample_ratio: int = 8):
        if filter not in ("ideal", "ideal_crop"):
            raise NotImplementedError(f"afldm_b200 provides the ideal shifters on device, not {filter!r}")
        if upsample_ratio is None or int(upsample_ratio) < 1:
            raise ValueError("upsample_ratio must be a positive integer")
        self.filter = filter
        self.upsample_ratio = int(upsample_ratio)

    def shift(self, img: torch.Tensor, ti: float, tj: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """img [B,C,H,W] fp32 CUDA -> (warped [B,C,H,W], mask [B,C,H,W]); rows move by ti, columns by tj."""
        warped, masks = self.shift_batch(img, [(ti, tj)])
        return warped[0], masks[0]

    def shift_batch(self, img: torch.Tensor, shifts: Sequence[Tuple[float, float]]) -> Tuple[torch.Tensor, torch.Tensor]:
        """All ``shifts`` of the same image in one launch pair: -> (warped [S,B,C,H,W], masks [S,B,C,H,W])."""
        if not img.is_cuda or img.dtype != torch.float32 or img.ndim != 4:
            raise _lib.AfldmError("ImageShifter: fp32 CUDA [B,C,H,W] tensor expected (no CPU fallback)")
        x = img.contiguous()
        b, c, h, w = x.shape
        ns = len(shifts)
        if ns == 0:
            raise ValueError("no shifts given")
        crop = self.filter == "ideal_crop"
        r = self.upsample_ratio
        my = np.stack([shift_matrix(h, float(ti), r, crop) for ti, _ in shifts]).astype(np.float32)
        mx = np.stack([shift_matrix(w, float(tj), r, crop) for _, tj in shifts]).astype(np.float32)
        my_d = torch.from_numpy(my).to(x.device)
        mx_d = torch.from_numpy(mx).to(x.device)
        # one plane list [S][B*C]: every shift reads the same image planes
        xs = x.view(1, b * c, h, w).expand(ns, b * c, h, w).contiguous() if ns > 1 else x.view(1, b * c, h, w)
        out = torch.empty((ns, b, c, h, w), dtype=torch.float32, device=x.device)
        L = _lib.lib()
        planes = ns * b * c
        need = L.afldm_plane_sep_transform_workspace_floats(planes, h, w)
        ws = ops.scratch(x.device, need)
        ops._run("plane_sep_transform", dict(elems=planes * h * w),
                 lambda: L.afldm_plane_sep_transform_f32(xs.data_ptr(), my_d.data_ptr(), mx_d.data_ptr(), out.data_ptr(),
                                                         ws.data_ptr(), need, planes, b * c, h, w, h, w,
                                                         ops._stream()), (xs, my_d, mx_d, out, ws))
        if crop:
            masks = torch.stack([gen_valid_mask((b, c, h, w), ti, tj) for ti, tj in shifts]).to(x.device)
        else:
            masks = torch.ones_like(out)
        return out, masks
