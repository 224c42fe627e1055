"""File helpers with the surface of /root/reference/afldm/io_utils.py:10-61 (``image_to_tensor``,
``save_gif_from_tensors``).  Plain host-side I/O (PIL), outside the measured path."""
from __future__ import annotations

import os

import numpy as np
import torch


def _load_rgb(file):
    import PIL.Image
    import PIL.ImageOps
    img = PIL.Image.open(file) if isinstance(file, (str, os.PathLike)) else file
    return PIL.ImageOps.exif_transpose(img).convert("RGB")


def image_to_tensor(files, resolution=(512, 512)) -> torch.Tensor:
    """io_utils.py:10-40: files (path, PIL image or list) -> [N,3,H,W] in [-1, 1]; bilinear resize to ``resolution``
    ((h, w)) unless it is None."""
    import PIL.Image
    if not isinstance(files, list):
        files = [files]
    out = []
    for f in files:
        img = _load_rgb(f)
        if resolution is not None:
            h, w = resolution
            img = img.resize((w, h), PIL.Image.BILINEAR)
        t = torch.from_numpy(np.asarray(img, dtype=np.float32) / 255.0).permute(2, 0, 1)
        out.append(((t - 0.5) / 0.5).unsqueeze(0))
    return torch.cat(out)


def save_gif_from_tensors(tensors, output_gif_path, duration=0.5, denorm=False):
    """io_utils.py:43-61: each tensor ([C,H,W], or [N,C,H,W] laid side by side) becomes one GIF frame."""
    import PIL.Image
    frames = []
    for t in tensors:
        t = t.detach().float().cpu()
        if denorm:
            t = (t + 1) / 2
        if t.ndim == 4:
            t = torch.cat(list(t), dim=-1)                  # 'n c h w -> c h (n w)'
        if t.shape[0] == 4:
            t = t[:3]
        t = torch.clamp(t, 0, 1)
        arr = (t.permute(1, 2, 0).numpy() * 255.0).astype(np.uint8)
        frames.append(PIL.Image.fromarray(arr.squeeze(-1) if arr.shape[-1] == 1 else arr))
    d = os.path.dirname(output_gif_path)
    if d:
        os.makedirs(d, exist_ok=True)
    frames[0].save(output_gif_path, save_all=True, append_images=frames[1:], duration=int(duration * 1000), loop=0)
