"""Oracle (test infrastructure): the fractional-shift equivariance measurement.

Follows /root/reference/afldm/shift_utils/shifters.py:31-49 (gen_valid_mask), :157-191
(ImageShifter 'ideal' / 'ideal_crop') and /root/reference/afldm/shift_utils/metrics.py:5-19.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import ideal_lpf as L


def gen_valid_mask(shape, ti, tj):
    _, _, h, w = shape
    if ti >= 0:
        i1, i2 = 0, int(np.ceil(ti))
    else:
        i1, i2 = int(np.floor(ti)), h
    if tj >= 0:
        j1, j2 = 0, int(np.ceil(tj))
    else:
        j1, j2 = int(np.floor(tj)), w
    mask = torch.ones(shape, dtype=torch.float32)
    mask[:, :, i1:i2, :] = 0
    mask[:, :, :, j1:j2] = 0
    return mask


def ideal_shift(img, ti, tj, ratio, crop=True):
    """ImageShifter('ideal_crop' | 'ideal', ratio).shift(img, ti, tj) -> (warped, mask)."""
    big = img if ratio == 1 else L.upsample_rfft(img, ratio)
    si, sj = int(np.round(ti * ratio)), int(np.round(tj * ratio))
    big = torch.roll(big, shifts=(si, sj), dims=(2, 3))
    if not crop:
        out = big[:, :, ::ratio, ::ratio]
        return out, torch.ones_like(out)
    big = big * gen_valid_mask(big.shape, si, sj).to(big.device)
    out = big[:, :, ::ratio, ::ratio]
    return out, gen_valid_mask(out.shape, ti, tj).to(out.device)


def mask_mse(a, b, mask):
    per = (a * mask - b * mask).square().sum((1, 2, 3)) / mask.sum((1, 2, 3))
    return per.mean()


def mask_psnr(a, b, mask):
    am, bm = a * mask, b * mask
    rng = torch.max(am.max(), bm.max()) - torch.min(am.min(), bm.min())
    return 10 * torch.log10(rng * rng / mask_mse(a, b, mask))


def psnr(a, b, i_max=None):
    if i_max is None:
        i_max = torch.max(a.max(), b.max()) - torch.min(a.min(), b.min())
    return 10 * torch.log10(i_max * i_max / F.mse_loss(a, b))


def _sinc(x: np.ndarray) -> np.ndarray:
    """af_libs/equivariance.py:23-26."""
    y = np.abs(x * np.pi)
    return np.where(y < 1e-30, 1.0, np.sin(y) / np.maximum(y, 1e-30))


def lanczos_shift(img: torch.Tensor, ti: float, tj: float, a: int = 3):
    """``ImageShifter('lanczos').shift`` (shifters.py:159-162) = ``apply_fractional_translation(img, tj / w, ti / h)``
    (af_libs/equivariance.py:68-103), restated step by step as the reference runs it: split the offset into integer and
    fractional parts, build the 2a-tap Lanczos filters at the fractional offset, normalise them, filter each axis with
    zero padding (``filter2d`` with padding [b, a] = a true convolution whose output is b + a samples longer than the
    image), crop, and paste the result at the integer offset; the mask is one where the whole filter support was inside."""
    x = img.double().numpy()
    n, c, h, w = x.shape
    tx, ty = np.float32((tj / w) * w), np.float32((ti / h) * h)
    ix, iy = int(np.floor(tx)), int(np.floor(ty))
    fx, fy = float(np.float32(tx - ix)), float(np.float32(ty - iy))
    b = a - 1
    z = np.zeros_like(x)
    zx0, zy0 = max(ix - b, 0), max(iy - b, 0)
    zx1, zy1 = min(ix + a, 0) + w, min(iy + a, 0) + h
    if zx0 < zx1 and zy0 < zy1:
        taps = np.arange(2 * a) - b
        f_x = _sinc(taps - fx) * _sinc((taps - fx) / a)
        f_y = _sinc(taps - fy) * _sinc((taps - fy) / a)
        f_x, f_y = f_x / f_x.sum(), f_y / f_y.sum()
        pad = b + a                                        # filter2d: caller's [b, a] + its own [fw // 2, (fw - 1) // 2]
        xp = np.zeros((n, c, h, w + 2 * pad))
        xp[..., pad:pad + w] = x
        y = np.zeros((n, c, h, w + pad))
        for p in range(w + pad):                           # true convolution: y[p] = sum_k f[k] xp[p + K - 1 - k]
            for k in range(2 * a):
                y[..., p] += f_x[k] * xp[..., p + 2 * a - 1 - k]
        yp = np.zeros((n, c, h + 2 * pad, w + pad))
        yp[:, :, pad:pad + h, :] = y
        y2 = np.zeros((n, c, h + pad, w + pad))
        for p in range(h + pad):
            for k in range(2 * a):
                y2[:, :, p, :] += f_y[k] * yp[:, :, p + 2 * a - 1 - k, :]
        y2 = y2[:, :, max(b - iy, 0): h + b + a + min(-iy - a, 0), max(b - ix, 0): w + b + a + min(-ix - a, 0)]
        z[:, :, zy0:zy1, zx0:zx1] = y2
    m = np.zeros((n, 1, h, w), dtype=np.float32)
    mx0, my0 = max(ix + a, 0), max(iy + a, 0)
    mx1, my1 = min(ix - b, 0) + w, min(iy - b, 0) + h
    if mx0 < mx1 and my0 < my1:
        m[:, :, my0:my1, mx0:mx1] = 1
    return torch.from_numpy(z.astype(np.float32)), torch.from_numpy(m)
