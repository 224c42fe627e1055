"""The part of /root/reference/afldm/shift_utils/flow_utils.py the shift measurement uses (:34-86): ``coords_grid``,
``bilinear_sample`` / ``flow_warp`` for UNIFORM flows (every pixel of a sample moves by the same vector - what
``ImageShifter`` and ``image_random_translate`` build, shifters.py:84-93, 200-205).  A uniform bilinear warp is
separable, so it runs as one matrix per axis on ``afldm_plane_sep_transform_f32``; general (per-pixel) flow fields
belong to the GMFlow interpolation application, which is outside this build."""
from __future__ import annotations

import numpy as np
import torch

from .shifters import _sep, bilinear_matrix


def coords_grid(b, h, w, homogeneous=False, device=None):
    """flow_utils.py:34-50: [B, 2 (x, y), H, W] pixel coordinates ([B, 3, H, W] with a ones plane if homogeneous)."""
    y, x = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
    stacks = [x, y] + ([torch.ones_like(x)] if homogeneous else [])
    grid = torch.stack(stacks, dim=0).float()[None].repeat(b, 1, 1, 1)
    return grid.to(device) if device is not None else grid


def _uniform(flow: torch.Tensor):
    """Per-sample (dx, dy) of a uniform flow field [B,2,H,W]; raises when the field varies over the image."""
    f0 = flow[:, :, :1, :1]
    if not bool((flow == f0).all()):
        raise NotImplementedError("flow_warp: only uniform (pure translation) flow fields run on the sm_100a path")
    return f0[:, 0, 0, 0].tolist(), f0[:, 1, 0, 0].tolist()


def flow_warp(feature, flow, mask=False, mode="bilinear", padding_mode="zeros"):
    """flow_utils.py:79-86: ``flow`` [B,2,H,W] in (row, column) order (the reference flips it to (x, y) :82); sample
    (i, j) of the output reads feature at (i + flow_row, j + flow_col), bilinear, zeros outside.  With ``mask=True``
    also returns the [B,H,W] bool map of positions that fall inside the image (:72-74)."""
    if mode != "bilinear" or padding_mode != "zeros":
        raise NotImplementedError("flow_warp: bilinear / zeros only")
    b, c, h, w = feature.shape
    assert flow.size(1) == 2
    rows, cols = _uniform(flow)
    outs, masks = [], []
    for s in range(b):
        (my, vr), (mx, vc) = bilinear_matrix(h, -rows[s]), bilinear_matrix(w, -cols[s])
        outs.append(_sep(feature[s:s + 1].contiguous(), my[None], mx[None])[0])
        masks.append(torch.from_numpy(np.outer(vr, vc)))
    out = torch.cat(outs, dim=0)
    if mask:
        return out, torch.stack(masks).to(feature.device)
    return out


def bilinear_sample(img, sample_coords, mode="bilinear", padding_mode="zeros", return_mask=False):
    """flow_utils.py:53-76 for a sampling grid that is a pure translation of the pixel grid."""
    if sample_coords.size(1) != 2:
        sample_coords = sample_coords.permute(0, 3, 1, 2)
    b, _, h, w = sample_coords.shape
    flow_xy = sample_coords - coords_grid(b, h, w, device=sample_coords.device)
    return flow_warp(img, torch.flip(flow_xy, (1,)), return_mask, mode, padding_mode)
