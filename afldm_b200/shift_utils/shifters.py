"""Fractional-shift warps of the equivariance measurement, on device.

Mirror of ``ImageShifter`` (/root/reference/afldm/shift_utils/shifters.py:135-206:
``ImageShifter(filter=None, upsample_ratio=None).shift(img, ti, tj) -> (warped, mask)`` with filter None = 'bilinear'
(the reference's default: a uniform bilinear ``flow_warp``, flow_utils.py:53-86), 'ideal', 'ideal_crop', 'fourier',
'fourier_crop'), of ``gen_valid_mask`` (:31-49), ``gen_random_offset`` (:52-76) and ``fourier_shift_batch`` (:101-132).
Every one of these warps is a uniform translation, i.e. linear and separable: one matrix per axis, applied by
``afldm_plane_sep_transform_f32``.  The reference up-samples the image r-fold with ``UpsampleRFFT(r)`` (64x the data at
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
from enum import Enum
from typing import Optional, Sequence, Tuple

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


FILTER_CHOICES = ['bilinear', 'lanczos', 'ideal', 'ideal_crop', 'fourier', 'fourier_crop']


def gen_random_offset(max_offset_i, max_offset_j, int_offset, int_stride, bs=1, min_offset_i=0, min_offset_j=0):
    """shifters.py:52-76: random (rows, columns) offsets, integer multiples of ``int_stride`` or uniform reals."""
    len_i, len_j = max_offset_i - min_offset_i, max_offset_j - min_offset_j
    if int_offset:
        ri, rj = int(len_i // int_stride), int(len_j // int_stride)
        oi = torch.randint(-ri, ri + 1, (bs,)).to(torch.float32) * int_stride
        oj = torch.randint(-rj, rj + 1, (bs,)).to(torch.float32) * int_stride
    else:
        oi = (torch.rand((bs,)) * 2 - 1) * len_i
        oj = (torch.rand((bs,)) * 2 - 1) * len_j
    return oi + min_offset_i, oj + min_offset_j


def bilinear_matrix(n: int, t: float) -> Tuple[np.ndarray, np.ndarray]:
    """One axis of the reference's default warp: ``flow_warp`` with the uniform backward flow -t
    (shifters.py:200-205 -> flow_utils.py:79-86 -> ``F.grid_sample(bilinear, zeros, align_corners=True)``).
    Output sample o reads position o - t.  The coordinate goes through the same fp32 round trip as the reference
    (normalise to [-1, 1] :65-66, un-normalise inside grid_sample) so that the interpolation weights - and the validity
    mask ``-1 <= grid <= 1`` (:72-74) - agree with it to the last bit of the coordinate.  Returns (matrix [n, n] fp64,
    valid [n] bool)."""
    f32 = np.float32
    pos = np.arange(n, dtype=np.float32) + f32(-t)                      # coords_grid + flow, fp32
    grid = f32(2) * pos / f32(n - 1) - f32(1)                           # bilinear_sample: normalise
    valid = (grid >= f32(-1)) & (grid <= f32(1))
    src = ((grid + f32(1)) / f32(2)) * f32(n - 1)                       # grid_sampler_unnormalize, align_corners=True
    i0 = np.floor(src)
    w1 = (src - i0).astype(np.float64)
    w0 = ((i0 + f32(1)) - src).astype(np.float64)
    m = np.zeros((n, n))
    for o in range(n):
        a = int(i0[o])
        if 0 <= a < n:
            m[o, a] += w0[o]
        if 0 <= a + 1 < n:
            m[o, a + 1] += w1[o]
    return m, valid


def lanczos_matrix(n: int, t: float, a: int = 3) -> Tuple[np.ndarray, np.ndarray]:
    """One axis of StyleGAN3's ``apply_fractional_translation`` (af_libs/equivariance.py:68-103, what
    ``ImageShifter('lanczos')`` calls with tx = tj / w, ty = ti / h, shifters.py:159-162): output sample X is
    sum_m x[m] L(X - m - t) / S over the 2a taps X - m - floor(t) in {-(a-1) .. a}, L(u) = sinc(u) sinc(u / a),
    S = the sum of the 2a tap values (the filter is normalised before it meets the border), zeros outside the image;
    columns outside [max(floor(t) - a + 1, 0), min(floor(t) + a, 0) + n) are zero.  The offset goes through the same
    fp32 value as the reference's ``torch.as_tensor(tx * W).to(float32)``.  Returns (matrix [n, n] fp64, valid [n] bool:
    the reference's mask, ones on [max(floor(t) + a, 0), min(floor(t) - a + 1, 0) + n))."""
    f32 = np.float32
    t32 = f32((t / n) * n)
    it = int(np.floor(t32))
    ft = f32(t32 - f32(it))
    b = a - 1
    taps = np.arange(-b, a + 1 - 0, dtype=np.float32)[:2 * a]                  # -(a-1) .. a
    u = (taps - ft).astype(np.float32)

    def sinc32(v):
        y = np.abs(v.astype(np.float32) * f32(np.pi))
        with np.errstate(invalid="ignore", divide="ignore"):
            z = np.sin(y) / np.maximum(y, f32(1e-30))
        return np.where(y < 1e-30, f32(1), z).astype(np.float32)

    f = (sinc32(u) * sinc32(u / f32(a))).astype(np.float32)
    f = (f / f.sum(dtype=np.float32)).astype(np.float64)
    m = np.zeros((n, n))
    z0, z1 = max(it - b, 0), min(it + a, 0) + n
    for X in range(z0, max(z0, z1)):
        for k in range(2 * a):
            src = X - it - int(taps[k])
            if 0 <= src < n:
                m[X, src] += f[k]
    valid = np.zeros(n, dtype=bool)
    m0, m1 = max(it + a, 0), min(it - b, 0) + n
    if m0 < m1:
        valid[m0:m1] = True
    return m, valid


def fourier_matrices(n: int, t: float) -> Tuple[np.ndarray, np.ndarray]:
    """One axis of ``fourier_shift_batch`` (shifters.py:101-132): multiplication of the DFT by exp(-2 pi i t f_k),
    f = fftfreq(n).  Returns the real and imaginary circulants (A_r, A_i): the reference keeps the REAL part of the 2-D
    inverse transform, which for a real image is  A_r x B_r^T - A_i x B_i^T  (A_i only carries the Nyquist bin)."""
    k = np.fft.fftfreq(n)
    ker = np.fft.ifft(np.exp(-2j * np.pi * t * k))
    idx = (np.arange(n)[:, None] - np.arange(n)[None, :]) % n
    return np.real(ker)[idx], np.imag(ker)[idx]


def _sep(x: torch.Tensor, my: np.ndarray, mx: np.ndarray) -> torch.Tensor:
    """x [B,C,H,W]; my [S,Ho,H], mx [S,Wo,W] -> [S,B,C,Ho,Wo]: y[s] = my[s] x mx[s]^T per plane, one launch pair."""
    b, c, h, w = x.shape
    ns, ho, wo = my.shape[0], my.shape[1], mx.shape[1]
    my_d = torch.from_numpy(np.ascontiguousarray(my, dtype=np.float32)).to(x.device)
    mx_d = torch.from_numpy(np.ascontiguousarray(mx, dtype=np.float32)).to(x.device)
    xs = x.view(1, b * c, h, w).expand(ns, b * c, h, w).contiguous() if ns > 1 else x.view(1, b * c, h, w)
    out = torch.empty((ns, b, c, ho, wo), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    planes = ns * b * c
    need = L.afldm_plane_sep_transform_workspace_floats(planes, h, wo)
    ws = ops.scratch(x.device, need)
    ops._run("plane_sep_transform", dict(elems=planes * ho * wo),
             lambda: L.afldm_plane_sep_transform_f32(xs.data_ptr(), my_d.data_ptr(), mx_d.data_ptr(), out.data_ptr(),
                                                     ws.data_ptr(), need, planes, b * c, h, w, ho, wo, ops._stream()),
             (xs, my_d, mx_d, out, ws))
    return out


def fourier_shift_batch(image: torch.Tensor, shift_x, shift_y, device="cuda") -> torch.Tensor:
    """shifters.py:101-132 (shift_x moves rows, shift_y columns - the reference indexes its frequency grid 'ij')."""
    image = image.to(device)
    n, c, h, w = image.shape
    ar, ai = fourier_matrices(h, float(shift_x))
    br, bi = fourier_matrices(w, float(shift_y))
    re = _sep(image.contiguous(), ar[None], br[None])[0]
    if np.abs(ai).max() * np.abs(bi).max() == 0.0:
        return re
    im = _sep(image.contiguous(), ai[None], bi[None])[0]
    return ops.nchw_view(ops.axpby(ops.nhwc(re), ops.nhwc(im), 1.0, -1.0)).contiguous()


class ImageShifter:
    """``ImageShifter(filter=None, upsample_ratio=None)`` (shifters.py:135-206): bilinear (default), lanczos, ideal,
    ideal_crop, fourier, fourier_crop - every filter is a pair of operator matrices applied by one launch pair."""

    class BgType(Enum):
        NO_BG = 0
        RANDN = 1
        FULL_COLOR = 2
        ORIGINAL_IMG = 3

    def __init__(self, filter: Optional[str] = None, upsample_ratio: Optional[int] = None):
        if filter is None:
            filter = "bilinear"                             # the reference's default (:143-144)
        assert filter in FILTER_CHOICES, f"Wrong filter type {filter}"
        if filter in ("ideal", "ideal_crop"):
            assert upsample_ratio is not None
            if int(upsample_ratio) < 1:
                raise ValueError("upsample_ratio must be a positive integer")
            self.upsample_ratio = int(upsample_ratio)
        self.filter = filter

    def shift(self, img: torch.Tensor, ti: float, tj: float) -> Tuple[torch.Tensor, torch.Tensor]:
        """img [B,C,H,W] fp32 CUDA -> (warped [B,C,H,W], mask); rows move by ti, columns by tj.  The mask is
        [B,C,H,W] for the ideal / fourier filters and [B,1,H,W] for the bilinear / lanczos warps, as in the reference."""
        warped, masks = self.shift_batch(img, [(ti, tj)])
        return warped[0], masks[0]

    def shift_batch(self, img: torch.Tensor, shifts: Sequence[Tuple[float, float]]) -> Tuple[torch.Tensor, torch.Tensor]:
        """All ``shifts`` of the same image in one launch pair: -> (warped [S,B,C,H,W], masks [S,B,C|1,H,W])."""
        if not img.is_cuda or img.dtype != torch.float32 or img.ndim != 4:
            raise _lib.AfldmError("ImageShifter: fp32 CUDA [B,C,H,W] tensor expected (no CPU fallback)")
        x = img.contiguous()
        b, c, h, w = x.shape
        shifts = [(float(ti), float(tj)) for ti, tj in shifts]
        if len(shifts) == 0:
            raise ValueError("no shifts given")
        if self.filter in ("ideal", "ideal_crop"):
            crop = self.filter == "ideal_crop"
            r = self.upsample_ratio
            my = np.stack([shift_matrix(h, ti, r, crop) for ti, _ in shifts])
            mx = np.stack([shift_matrix(w, tj, r, crop) for _, tj in shifts])
            out = _sep(x, my, mx)
            if crop:
                masks = torch.stack([gen_valid_mask((b, c, h, w), ti, tj) for ti, tj in shifts]).to(x.device)
            else:
                masks = torch.ones_like(out)
            return out, masks
        if self.filter in ("fourier", "fourier_crop"):
            outs = torch.stack([fourier_shift_batch(x, ti, tj, x.device) for ti, tj in shifts])
            if self.filter == "fourier":
                return outs, torch.ones_like(outs)
            masks = torch.stack([gen_valid_mask((b, c, h, w), ti, tj) for ti, tj in shifts]).to(x.device)
            return outs * masks, masks
        # bilinear (default): uniform flow warp, zeros outside, mask = sampled position inside the image;
        # lanczos: the 6-tap windowed-sinc translation of StyleGAN3's equivariance metrics, mask = full filter support
        axis = lanczos_matrix if self.filter == "lanczos" else bilinear_matrix
        rows = [axis(h, ti) for ti, _ in shifts]
        cols = [axis(w, tj) for _, tj in shifts]
        out = _sep(x, np.stack([m for m, _ in rows]), np.stack([m for m, _ in cols]))
        masks = torch.stack([torch.from_numpy(np.outer(vr, vc).astype(np.float32)) for (_, vr), (_, vc) in zip(rows, cols)])
        masks = masks.view(len(shifts), 1, 1, h, w).expand(len(shifts), b, 1, h, w).contiguous().to(x.device)
        return out, masks

    def translate_with_occ_bg(self, img: torch.Tensor, ti: float, tj: float, bg_type: "ImageShifter.BgType",
                              mask: Optional[torch.Tensor] = None, return_mask: bool = False):
        """shifters.py:208-236: the warp composited over a background where the mask is 0."""
        background = None
        if bg_type == ImageShifter.BgType.RANDN:
            background = torch.randn_like(img)
        elif bg_type == ImageShifter.BgType.FULL_COLOR:
            n, c = img.shape[0:2]
            background = (torch.rand((n, c, 1, 1)) * 2 - 1).to(device=img.device, dtype=img.dtype)
        elif bg_type == ImageShifter.BgType.ORIGINAL_IMG:
            background = img
        elif bg_type != ImageShifter.BgType.NO_BG:
            raise ValueError(f"No such background type {bg_type} in image shifter")
        warped, translate_mask = self.shift(img, ti, tj)
        if mask is None:
            mask = translate_mask
        if background is not None:
            warped = warped * mask + background * (1 - mask)
        return (warped, mask) if return_mask else warped
