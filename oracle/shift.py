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
