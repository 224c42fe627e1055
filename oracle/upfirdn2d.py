"""Oracle (test infrastructure): upfirdn2d in its slow reference form.

Follows /root/reference/afldm/af_libs/torch_utils/ops/upfirdn2d.py:70-116 (setup_filter),
:166-211 (_upfirdn2d_ref) and :313-348 (upsample2d padding rule).  Direct gather form in
numpy fp64 so that it is an independent statement of: zero-insert (up), pad/crop, true
convolution with f (correlation if flip_filter), decimate (down), times gain.
"""
from __future__ import annotations

import numpy as np


def setup_filter(taps, normalize=True, flip_filter=False, gain=1.0):
    f = np.asarray(taps, dtype=np.float32)
    if f.ndim == 0:
        f = f[None]
    separable = f.ndim == 1 and f.size >= 8
    if f.ndim == 1 and not separable:
        f = np.outer(f, f)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f[tuple(slice(None, None, -1) for _ in range(f.ndim))]
    return (f * gain ** (f.ndim / 2)).astype(np.float32)


def upfirdn2d(x, f, up=1, down=1, padding=(0, 0, 0, 0), flip_filter=False, gain=1.0):
    """x: (B,C,H,W) array; f: (fh,fw) or (taps,) separable. padding = (x0, x1, y0, y1)."""
    x = np.asarray(x, dtype=np.float64)
    f = np.asarray(f, dtype=np.float64)
    if f.ndim == 1:
        f = np.outer(f, f)
    fh, fw = f.shape
    px0, px1, py0, py1 = padding
    b, c, h, w = x.shape
    z = np.zeros((b, c, h * up, w * up))
    z[:, :, ::up, ::up] = x
    z = np.pad(z, ((0, 0), (0, 0), (max(py0, 0), max(py1, 0)), (max(px0, 0), max(px1, 0))))
    z = z[:, :, max(-py0, 0): z.shape[2] - max(-py1, 0), max(-px0, 0): z.shape[3] - max(-px1, 0)]
    k = f * gain
    if not flip_filter:
        k = k[::-1, ::-1]          # true convolution == correlation with the flipped kernel
    oh, ow = z.shape[2] - fh + 1, z.shape[3] - fw + 1
    out = np.zeros((b, c, oh, ow))
    for i in range(fh):
        for j in range(fw):
            out += k[i, j] * z[:, :, i:i + oh, j:j + ow]
    return out[:, :, ::down, ::down]


def upsample2d(x, f, up=2, gain=1.0):
    fa = np.asarray(f)
    fh = fa.shape[0]
    fw = fa.shape[-1]
    p = ((fw + up - 1) // 2, (fw - up) // 2, (fh + up - 1) // 2, (fh - up) // 2)
    return upfirdn2d(x, f, up=up, padding=p, gain=gain * up * up)


def downsample2d(x, f, down=2, gain=1.0):
    fa = np.asarray(f)
    fh = fa.shape[0]
    fw = fa.shape[-1]
    p = ((fw - down + 1) // 2, (fw - down) // 2, (fh - down + 1) // 2, (fh - down) // 2)
    return upfirdn2d(x, f, down=down, padding=p, gain=gain)
