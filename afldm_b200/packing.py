"""Weight packing for the implicit-GEMM kernels, cached on the owning nn.Module.

``nn.Conv2d`` keeps its diffusers-compatible ``weight`` [Cout,Cin,kh,kw] / ``bias`` Parameters (so
checkpoints load unchanged); the kernels read a tap-major copy [Cout][kh*kw][Cin] that is rebuilt
whenever the Parameter's storage or version changes.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import ops


def _key(ts: Sequence[Optional[torch.Tensor]]):
    return tuple(None if t is None else (t.data_ptr(), t._version, str(t.device)) for t in ts)


def conv_params(m: nn.Module) -> Tuple[torch.Tensor, Optional[torch.Tensor], int]:
    """(packed weight, bias, ksize) of an nn.Conv2d (k in {1,3}) or nn.Linear (as k = 1)."""
    key = _key((m.weight, m.bias))
    cache = getattr(m, "_afldm_pack", None)
    if cache is None or cache[0] != key:
        w = m.weight
        if w.ndim == 4:
            k = w.shape[-1]
            if w.shape[-2] != k or k not in (1, 3):
                raise NotImplementedError(f"conv kernel {tuple(w.shape[-2:])} not supported")
        else:
            k = 1
        bias = None if m.bias is None else m.bias.detach().contiguous()
        cache = (key, (ops.pack_conv_weight(w), bias, k))
        object.__setattr__(m, "_afldm_pack", cache)
    return cache[1]


def conv_params_f16(m: nn.Module) -> Tuple[torch.Tensor, Optional[torch.Tensor], int]:
    """``conv_params`` with the packed weight rounded to fp16 (the B operand of ``ops.conv2d`` on an fp16 activation:
    tcgen05.mma.kind::f16 keeps the same 11 significant bits TF32 does, in half the bytes)."""
    key = _key((m.weight, m.bias))
    cache = getattr(m, "_afldm_pack_f16", None)
    if cache is None or cache[0] != key:
        w, bias, k = conv_params(m)
        cache = (key, (w.to(torch.float16).contiguous(), bias, k))
        object.__setattr__(m, "_afldm_pack_f16", cache)
    return cache[1]


def fused_linear_params(owner: nn.Module, tag: str, layers: Sequence[nn.Module]):
    """Row-concatenated weight / bias of several nn.Linear (or 1x1 conv) layers sharing an input:
    q|k|v of an attention block, or every resnet's time_emb_proj of a UNet."""
    key = _key([t for m in layers for t in (m.weight, m.bias)])
    name = "_afldm_fused_" + tag
    cache = getattr(owner, name, None)
    if cache is None or cache[0] != key:
        w = torch.cat([m.weight.detach().reshape(m.weight.shape[0], -1) for m in layers], dim=0).contiguous()
        # bias-free projections (SD-1.5 transformer attention: to_q / to_k / to_v have no bias)
        b = None if all(m.bias is None for m in layers) else torch.cat(
            [m.bias.detach() if m.bias is not None else m.weight.new_zeros(m.weight.shape[0]) for m in layers], dim=0).contiguous()
        cache = (key, (w, b))
        object.__setattr__(owner, name, cache)
    return cache[1]


def fused_linear_params_f16(owner: nn.Module, tag: str, layers: Sequence[nn.Module]):
    """``fused_linear_params`` with the weight rounded to fp16 (B operand of the kind::f16 projection)."""
    key = _key([t for m in layers for t in (m.weight, m.bias)])
    name = "_afldm_fused_f16_" + tag
    cache = getattr(owner, name, None)
    if cache is None or cache[0] != key:
        w, b = fused_linear_params(owner, tag, layers)
        cache = (key, (w.to(torch.float16).contiguous(), b))
        object.__setattr__(owner, name, cache)
    return cache[1]


def conv_params_padded(m: nn.Module, cin_pad: int) -> Tuple[torch.Tensor, Optional[torch.Tensor], int]:
    """``conv_params`` with the input channels zero-padded to ``cin_pad`` (conv_in: 4 latent channels -> one 32-channel
    TMA chunk, so the layer runs on the tensor-core path; the extra channels multiply zero weights)."""
    key = _key((m.weight, m.bias)) + (cin_pad,)
    cache = getattr(m, "_afldm_pack_pad", None)
    if cache is None or cache[0] != key:
        w = m.weight.detach()
        co, ci, kh, kw = w.shape
        wp = torch.zeros((co, cin_pad, kh, kw), dtype=w.dtype, device=w.device)
        wp[:, :ci] = w
        bias = None if m.bias is None else m.bias.detach().contiguous()
        cache = (key, (ops.pack_conv_weight(wp), bias, kh))
        object.__setattr__(m, "_afldm_pack_pad", cache)
    return cache[1]
