"""StyleGAN3 ``upfirdn2d`` Python surface (/root/reference/afldm/af_libs/torch_utils/ops/upfirdn2d.py:70-387:
``setup_filter``, ``upfirdn2d``, ``filter2d``, ``upsample2d``, ``downsample2d``) on the sm_100a gather kernel
``afldm_upfirdn2d_f32``.  Same argument meaning (scaling int or [x, y]; padding int, [x, y] or [x0, x1, y0, y1];
``flip_filter=False`` = true convolution; 1-D ``f`` = separable) and the same output-size rule
``(in * up + pad0 + pad1 - f + down) // down`` (upfirdn2d.cpp:35-36).  ``impl`` is accepted for signature
compatibility; there is one implementation here (CUDA) and no CPU fallback."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib, ops


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(v, int) for v in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(v, int) for v in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in (1, 2)
    return int(f.shape[-1]), int(f.shape[0])


def setup_filter(f, device=torch.device("cpu"), normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:70-116."""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in (0, 1, 2) and f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]
    if separable is None:
        separable = f.ndim == 1 and f.numel() >= 8
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f.to(device=device)


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """upfirdn2d.py:118-162: zero-insert (up), pad / crop, FIR filter, decimate (down), times gain."""
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    if not x.is_cuda:
        raise _lib.AfldmError("upfirdn2d: CUDA tensor expected (afldm_b200 has no CPU implementation)")
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    if upx != upy or downx != downy:
        raise NotImplementedError("upfirdn2d: equal x / y scaling factors only")
    px0, px1, py0, py1 = _parse_padding(padding)
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
    f = f.to(device=x.device, dtype=torch.float32)
    # a separable 1-D filter applies along both axes: its outer product, with the gain applied once
    return ops.upfirdn2d(x.contiguous().float(), f, up=upx, down=downx, padding=(px0, px1, py0, py1),
                         flip_filter=flip_filter, gain=float(gain))


def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """upfirdn2d.py:277-310: same-size output by default."""
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """upfirdn2d.py:313-348: output a multiple of the input, gain compensated by up^2."""
    upx, upy = _parse_scaling(up)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl="cuda"):
    """upfirdn2d.py:352-387."""
    downx, downy = _parse_scaling(down)
    px0, px1, py0, py1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)
