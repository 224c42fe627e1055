"""Torch-tensor front end of the C-ABI ops (``include/afldm_b200.h``).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every op below launches
hand-written sm_100a kernels from ``libafldm_b200.so`` on ``torch.cuda.current_stream()``.
Activations are fp32 and physically NHWC: functions take / return tensors of shape
``[B, H, W, C]`` (contiguous).  ``nhwc()`` / ``nchw_view()`` convert from and to the logical
``[B, C, H, W]`` shape the reference's modules use (free when the tensor is ``channels_last``).

There is no fallback path: on a machine without the library or without a GPU these raise.
"""
from __future__ import annotations

from typing import Optional, Tuple

import os

import torch

from . import _lib

ACT = {"identity": 0, None: 0, "silu": 1}
CONV_ALGO = {"simt": 0, "tf32": 1}

_default_conv_algo = "simt"
FUSE_GN_PROLOGUE = os.environ.get("AFLDM_FUSE_GN", "1") == "1"


def set_default_conv_algo(name: str) -> None:
    """'simt' = exact fp32 FMA; 'tf32' = tcgen05 tensor cores (SIMT where a shape does not fit)."""
    global _default_conv_algo
    if name not in CONV_ALGO:
        raise ValueError(name)
    _default_conv_algo = name


def default_conv_algo() -> str:
    return _default_conv_algo


_recorder = None


def record_to(records) -> None:
    """Profiling hook (bench.py): while a list is installed, every op appends
    ``(name, meta, thunk, keepalive)`` where ``thunk()`` re-issues the identical C call."""
    global _recorder
    _recorder = records


def _run(name: str, meta: dict, fn, keep=()) -> None:
    _lib.check(fn(), name)
    if _recorder is not None:
        _recorder.append((name, meta, fn, keep))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.AfldmError(f"{name}: afldm_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dtype != torch.float32:
        raise _lib.AfldmError(f"{name}: fp32 expected, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.AfldmError(f"{name}: contiguous tensor expected")
    return t


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------- scratch
_scratch = {}
_scratch_slot = 0


class scratch_slot:
    """Context manager: ops issued inside use scratch buffer ``slot`` (one per concurrently running stream)."""

    def __init__(self, slot: int):
        self.slot = slot

    def __enter__(self):
        global _scratch_slot
        self.prev, _scratch_slot = _scratch_slot, self.slot

    def __exit__(self, *exc):
        global _scratch_slot
        _scratch_slot = self.prev
_scratch_retired = []   # outgrown buffers stay alive: captured CUDA graphs may still point at them


def scratch(device: torch.device, nfloats: int) -> torch.Tensor:
    """One grow-only fp32 scratch buffer per (device, slot) (split-K partials, large-plane intermediates).
    Ops on one stream run in order, so consecutive users may share it; code that runs concurrently on a second
    stream (the second half-batch branch of a captured step, a resnet's conv_shortcut side stream) selects its own
    buffer with ``scratch_slot(k)``.  RESTRICTION: work issued by the USER on two streams of one device at the same
    time (say a UNet step on one stream and a VAE decode on another) must wrap one of them in ``scratch_slot(k)``,
    k >= 4, as well - the library keeps no per-stream state ("Threading / streams")."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), _scratch_slot)
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nfloats:
        if torch.cuda.is_current_stream_capturing():
            raise _lib.AfldmError("scratch would grow during CUDA-graph capture: run one eager warm-up first")
        if buf is not None:
            _scratch_retired.append(buf)
        buf = torch.empty(max(nfloats, 1 << 20), dtype=torch.float32, device=device)
        _scratch[key] = buf
    return buf


_side_streams = {}
# conv_shortcut of a resnet on a side stream (a parallel branch of the captured step graph): it depends on the block
# input only and is needed by conv2's epilogue, so it runs under norm1 / activation / conv1 instead of between them
SHORTCUT_SIDE_STREAM = os.environ.get("AFLDM_SC_SIDE", "1") == "1"
SIDE_SCRATCH_SLOT = 2


def side_scratch_slot() -> int:
    """Scratch slot of the conv_shortcut side stream, derived from the caller's slot: two concurrently running
    branches (``scratch_slot(0)`` / ``scratch_slot(1)``) get the distinct side buffers 2 / 3."""
    return SIDE_SCRATCH_SLOT + _scratch_slot


def side_stream(dev: torch.device) -> "torch.cuda.Stream":
    """One auxiliary stream per (device, calling stream) for work that is independent of the main dependency chain."""
    idx = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    key = (idx, torch.cuda.current_stream(idx).cuda_stream)
    st = _side_streams.get(key)
    if st is None:
        st = _side_streams[key] = torch.cuda.Stream(device=idx)
    return st


_deferred_join = None


def defer_join(main: "torch.cuda.Stream", side: "torch.cuda.Stream") -> None:
    """``main`` must wait for ``side`` before the NEXT op that consumes a time-embedding projection row (``conv2d`` with
    ``row_add``, ``linear_rows``) - not earlier: the time-embedding branch (~70 us on its side stream) then also runs
    under the first resnet's GroupNorm + filtered activation instead of stalling the main stream right after conv_in."""
    global _deferred_join
    _deferred_join = (main, side)


def join_deferred() -> None:
    global _deferred_join
    if _deferred_join is not None:
        main, side = _deferred_join
        _deferred_join = None
        main.wait_stream(side)


# ------------------------------------------------------------------------- layout
def nhwc(x: torch.Tensor) -> torch.Tensor:
    """Logical [B,C,H,W] -> physical [B,H,W,C] contiguous (zero-copy for channels_last input)."""
    v = x.permute(0, 2, 3, 1)
    if v.is_contiguous():
        return _carry_gn(x, v)
    x = _chk(x.contiguous(), "x")
    b, c, h, w = x.shape
    y = torch.empty((b, h, w, c), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    _run("nchw_to_nhwc", dict(elems=x.numel()),
         lambda: L.afldm_nchw_to_nhwc_f32(x.data_ptr(), y.data_ptr(), b, c, h * w, _stream()), (x, y))
    return y


def nchw_view(y: torch.Tensor) -> torch.Tensor:
    """Physical [B,H,W,C] -> logical [B,C,H,W] (channels_last strides, zero-copy)."""
    return _carry_gn(y, y.permute(0, 3, 1, 2))


def _carry_gn(src: torch.Tensor, dst: torch.Tensor) -> torch.Tensor:
    """Views of one buffer share the GroupNorm partial sums its producing kernel emitted."""
    for name in ("_afldm_gn", "_afldm_gn2"):
        st = getattr(src, name, None)
        if st is not None:
            setattr(dst, name, st)
    return dst


def to_nchw_contiguous(y: torch.Tensor) -> torch.Tensor:
    """Physical [B,H,W,C] -> contiguous [B,C,H,W] (pipeline boundary)."""
    _chk(y, "y")
    b, h, w, c = y.shape
    out = torch.empty((b, c, h, w), dtype=torch.float32, device=y.device)
    L = _lib.lib()
    _run("nhwc_to_nchw", dict(elems=y.numel()),
         lambda: L.afldm_nhwc_to_nchw_f32(y.data_ptr(), out.data_ptr(), b, c, h * w, _stream()), (y, out))
    return out


# ------------------------------------------------------------------------- ideal resampling
GENERAL_PLANES = os.environ.get("AFLDM_GENERAL_PLANES", "1") == "1"
_UNSUPPORTED = (-1, -3)        # AFLDM_E_SHAPE / AFLDM_E_NOKERNEL: outside the fused kernels' family, nothing launched


def _general_resample(x: torch.Tensor, mode: str, act: str = "identity", scale: Optional[torch.Tensor] = None,
                      shift: Optional[torch.Tensor] = None) -> torch.Tensor:
    """The three ideal operators for ANY square plane size (n > 128, not a power of two, channel counts outside the
    fused kernels' groups): the same operator matrices U (2n x n) and D (n x 2n) applied per plane by
    ``afldm_plane_sep_transform_f32``, with the activation as an ``afldm_affine_act_f32`` pass over the 4x tensor.
    Like the reference it materialises the up-sampled tensor; unlike the fused kernels it is not a hot-path form
    (the UNet / VAE configurations of SURVEY.md 8(d) never reach it).  NHWC in, NHWC out."""
    from .af_libs import ideal_lpf as il
    b, h, w, c = x.shape
    if h != w:
        raise _lib.AfldmError(f"{mode}: square planes expected (the reference builds its mask from the last dimension "
                              f"only, ideal_lpf.py:81), got {h} x {w}")
    if scale is not None:                                    # GroupNorm affine: commutes with U (rows of U sum to 1),
        x = affine_act(x.view(b, h * w, c), scale, shift, "identity").view(b, h, w, c)   # applied first for clarity
    t = to_nchw_contiguous(x)
    if mode in ("fact", "up2"):
        u = il.upsample_matrix(h, 2)
        t = il.sep_transform(t, u, u)
    if mode == "fact" and act != "identity":
        flat = t.view(1, -1, 1024) if t.numel() % 1024 == 0 else t.view(1, -1, 4) if t.numel() % 4 == 0 else t.view(1, -1, 1)
        affine_act(flat, None, None, act, out=flat)
    if mode in ("fact", "down2"):
        n2 = t.shape[-1]
        d = il.filter_matrix(n2, 0.5, 0.0)[::2, :]
        t = il.sep_transform(t, d, d)
    return nhwc(t)


def filtered_act(x: torch.Tensor, scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None,
                 act: str = "silu", out: Optional[torch.Tensor] = None, out_half: bool = False) -> torch.Tensor:
    """WarpedNonlinearity (af_blocks.py:19-28) on NHWC x, optional folded GroupNorm affine.  ``out_half``: the result
    MAY be fp16 (a tensor-core convolution is its only consumer); fp32 where the library has no fp16 store."""
    _chk(x, "x")
    b, h, w, c = x.shape
    L = _lib.lib()
    need = L.afldm_resample_workspace_floats(0, b, h, w, c)
    ws = scratch(x.device, need) if need else None
    if out_half and out is None:
        outh = torch.empty(x.shape, dtype=torch.float16, device=x.device)

        def call_h():
            return L.afldm_filtered_act_f16out(x.data_ptr(), outh.data_ptr(), b, h, w, c, ACT[act],
                                               _ptr(scale), _ptr(shift), _ptr(ws), need, _stream())

        code = call_h()
        if code not in _UNSUPPORTED:
            _lib.check(code, "filtered_act_f16out")
            if _recorder is not None:
                _recorder.append(("filtered_act", dict(B=b, N=h, C=c, elems=x.numel(), f16out=1), call_h,
                                  (x, outh, scale, shift, ws)))
            return outh
    dst = torch.empty_like(x) if out is None else out

    def call():
        return L.afldm_filtered_act_f32(x.data_ptr(), dst.data_ptr(), b, h, w, c, ACT[act],
                                        _ptr(scale), _ptr(shift), _ptr(ws), need, _stream())

    code = call()
    if code in _UNSUPPORTED and GENERAL_PLANES and h == w:
        res = _general_resample(x, "fact", act, scale, shift)
        return res if out is None else out.copy_(res)
    _lib.check(code, "filtered_act")
    if _recorder is not None:
        _recorder.append(("filtered_act", dict(B=b, N=h, C=c, elems=x.numel()), call, (x, dst, scale, shift, ws)))
    return dst


FUSE_CONCAT = os.environ.get("AFLDM_FUSE_CONCAT", "1") == "1"


def conv_f16_supported(b: int, h: int, w: int, cin: int, cout: int, ksize: int = 3) -> bool:
    """Whether ``conv2d`` will accept this layer with fp16 operands (TF32 class, switch on, and the tcgen05 plan of
    csrc/conv_tc.cu covers the shape with 64-channel stages - ``afldm_conv2d_supported``, a host-side query).  The
    producer of the activation asks this BEFORE it decides to store fp16."""
    if not (F16_CONV and _default_conv_algo == "tf32"):
        return False
    return bool(_lib.lib().afldm_conv2d_supported(b, h, w, cin, cout, ksize, 2))


def filtered_act_groupnorm_cat(a: torch.Tensor, b: torch.Tensor, groups: int, eps: float, gamma: Optional[torch.Tensor],
                               beta: Optional[torch.Tensor], act: str = "silu", out_half: bool = False) -> Optional[torch.Tensor]:
    """act-filtered GroupNorm(torch.cat([a, b], channels)) on NHWC a, b WITHOUT materialising the concat (norm1 of an
    up-block resnet).  Needs the GroupNorm partial sums of both producers; returns None when this form does not
    apply (missing sums, planes > 32, channel groups straddling the sources) - the caller then concatenates."""
    ga, gb = getattr(a, "_afldm_gn", None), getattr(b, "_afldm_gn", None)
    if not (FUSE_CONCAT and FUSE_GN_PROLOGUE) or ga is None or gb is None:
        return None
    _chk(a, "a")
    _chk(b, "b")
    bsz, h, w, ca = a.shape
    cb = b.shape[-1]
    if b.shape[:-1] != a.shape[:-1] or h != w or h > 32 or ga[2] != ca or gb[2] != cb:
        return None
    L = _lib.lib()
    out = None

    def call():
        fn = L.afldm_filtered_act_gn_cat_f16out if out.dtype == torch.float16 else L.afldm_filtered_act_gn_cat_f32
        return fn(a.data_ptr(), b.data_ptr(), out.data_ptr(), bsz, h, w, ca, cb, ACT[act],
                  ga[0].data_ptr(), ga[1], gb[0].data_ptr(), gb[1], groups, float(eps),
                  _ptr(gamma), _ptr(beta), _stream())

    code = -3
    if out_half:          # fp16 result for a tensor-core consumer; -3 = this plane size only has the fp32 kernel
        out = torch.empty((bsz, h, w, ca + cb), dtype=torch.float16, device=a.device)
        code = call()
    if code == -3:
        out = torch.empty((bsz, h, w, ca + cb), dtype=torch.float32, device=a.device)
        code = call()
    if code == -3:
        return None
    _lib.check(code, "filtered_act_gn_cat")
    if _recorder is not None:
        _recorder.append(("filtered_act", dict(B=bsz, N=h, C=ca + cb, elems=out.numel(), fused_gn=1, cat=1), call,
                          (a, b, out, ga[0], gb[0], gamma, beta)))
    return out


def conv2d_cat(a: torch.Tensor, b: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], ksize: int,
               gn_stats: bool = False) -> Optional[torch.Tensor]:
    """conv2d(torch.cat([a, b], channels)) on the tensor-core path without materialising the concat (conv_shortcut of
    an up-block resnet).  None when unavailable (exact-fp32 class, shapes outside the tcgen05 family)."""
    if not FUSE_CONCAT or _default_conv_algo != "tf32":
        return None
    _chk(w_packed, "w_packed")
    bsz, h, w_, ca = a.shape
    cb = b.shape[-1]
    cout = w_packed.shape[0]
    if b.shape[:-1] != a.shape[:-1] or w_packed.shape[-1] != ca + cb:
        return None
    pa, pb = _pitch(a), _pitch(b)
    out = torch.empty((bsz, h, w_, cout), dtype=torch.float32, device=a.device)
    L = _lib.lib()
    need = L.afldm_conv2d_workspace_floats(bsz, h, w_, ca + cb, cout, ksize, 1)
    ws = scratch(a.device, need) if need else None
    slots = L.afldm_conv2d_gn_slots(bsz, h, w_, ca + cb, cout, ksize, 1) if gn_stats else 0
    gn = torch.empty((bsz, slots, cout, 2), dtype=torch.float32, device=a.device) if slots else None

    def call():
        return L.afldm_conv2d_cat_f32(a.data_ptr(), pa, ca, b.data_ptr(), pb, cb, w_packed.data_ptr(), _ptr(bias), None, 0,
                                      None, 0, out.data_ptr(), cout, bsz, h, w_, cout, ksize, _ptr(ws), need, _ptr(gn),
                                      _stream())

    code = call()
    if code == -3:
        return None
    _lib.check(code, "conv2d_cat")
    if _recorder is not None:
        _recorder.append(("conv2d_tf32", dict(B=bsz, H=h, W=w_, Cin=ca + cb, Cout=cout, k=ksize, cat=1,
                                              flops=2.0 * bsz * h * w_ * cout * (ca + cb) * ksize * ksize), call,
                          (a, b, w_packed, bias, out, ws, gn)))
    if gn is not None:
        out._afldm_gn = (gn, slots, cout)
    return out


def filtered_act_groupnorm(x: torch.Tensor, groups: int, eps: float, gamma: Optional[torch.Tensor],
                           beta: Optional[torch.Tensor], act: str = "silu", out_half: bool = False) -> torch.Tensor:
    """act-filtered GroupNorm(x) on NHWC x: statistics pass (or finalize of the producer's partial sums) +
    filtered activation.  ``FUSE_GN_PROLOGUE`` selects the one-launch variant that finalises the statistics
    in the resampling kernel's prologue (afldm_filtered_act_gn_f32): one warp per touched group adds the producer's
    fp32 partial sums.  It costs the kernel ~2 us and removes a ~3 us finalize launch plus its gap (B200: +1 % on
    the step); ``AFLDM_FUSE_GN=0`` selects the two-launch form."""
    _chk(x, "x")
    b, h, w, c = x.shape
    one, two = getattr(x, "_afldm_gn", None), getattr(x, "_afldm_gn2", None)
    if FUSE_GN_PROLOGUE and (one is not None or two is not None) and h <= 32 and h == w:
        (pa, sa, ca), (pb, sb, cb) = (one, (None, 0, 0)) if one is not None else two
        if ca + cb == c:
            L = _lib.lib()
            if out_half:
                # ``out_half``: fp16 result (consumed by a tensor-core convolution only); falls through to fp32 where
                # the library has no fp16 store for this plane size (-3)
                outh = torch.empty(x.shape, dtype=torch.float16, device=x.device)

                def call_h():
                    return L.afldm_filtered_act_gn_f16out(x.data_ptr(), outh.data_ptr(), b, h, w, c, ACT[act],
                                                          pa.data_ptr(), sa, ca, _ptr(pb), sb, cb, groups, float(eps),
                                                          _ptr(gamma), _ptr(beta), _stream())

                code = call_h()
                if code not in _UNSUPPORTED:
                    _lib.check(code, "filtered_act_gn_f16out")
                    if _recorder is not None:
                        _recorder.append(("filtered_act", dict(B=b, N=h, C=c, elems=x.numel(), fused_gn=1, f16out=1),
                                          call_h, (x, outh, pa, pb, gamma, beta)))
                    return outh
            out = torch.empty_like(x)

            def call():
                return L.afldm_filtered_act_gn_f32(x.data_ptr(), out.data_ptr(), b, h, w, c, ACT[act], pa.data_ptr(),
                                                   sa, ca, _ptr(pb), sb, cb, groups, float(eps), _ptr(gamma),
                                                   _ptr(beta), _stream())

            code = call()
            if code not in _UNSUPPORTED:          # outside the fused family: statistics + the general-plane form below
                _lib.check(code, "filtered_act_gn")
                if _recorder is not None:
                    _recorder.append(("filtered_act", dict(B=b, N=h, C=c, elems=x.numel(), fused_gn=1), call,
                                      (x, out, pa, pb, gamma, beta)))
                return out
    scale, shift = groupnorm_affine(x, groups, eps, gamma, beta)
    return filtered_act(x, scale, shift, act=act, out_half=out_half)


def up2_ideal(x: torch.Tensor, scale: Optional[torch.Tensor] = None,
              shift: Optional[torch.Tensor] = None, out_half: bool = False) -> torch.Tensor:
    """UpsampleRFFT(2) (ideal_lpf.py:148-158) on NHWC x.  ``out_half``: the result MAY be fp16 (planes up to 16 x 16;
    for a tensor-core convolution as the only consumer)."""
    _chk(x, "x")
    b, h, w, c = x.shape
    if out_half and scale is None and h == w:
        outh = torch.empty((b, 2 * h, 2 * w, c), dtype=torch.float16, device=x.device)
        Lh = _lib.lib()
        needh = Lh.afldm_resample_workspace_floats(1, b, h, w, c)
        wsh = scratch(x.device, needh) if needh else None

        def call_h():
            return Lh.afldm_up2_ideal_f16out(x.data_ptr(), outh.data_ptr(), b, h, w, c, _ptr(wsh), needh, _stream())

        code = call_h()
        if code not in _UNSUPPORTED:
            _lib.check(code, "up2_ideal_f16out")
            if _recorder is not None:
                _recorder.append(("up2_ideal", dict(B=b, N=h, C=c, elems=x.numel(), f16out=1), call_h, (x, outh, wsh)))
            return outh
    out = torch.empty((b, 2 * h, 2 * w, c), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    need = L.afldm_resample_workspace_floats(1, b, h, w, c)
    ws = scratch(x.device, need) if need else None

    def call():
        return L.afldm_up2_ideal_f32(x.data_ptr(), out.data_ptr(), b, h, w, c, _ptr(scale), _ptr(shift),
                                     _ptr(ws), need, _stream())

    code = call()
    if code in _UNSUPPORTED and GENERAL_PLANES and h == w:
        return _general_resample(x, "up2", "identity", scale, shift)
    _lib.check(code, "up2_ideal")
    if _recorder is not None:
        _recorder.append(("up2_ideal", dict(B=b, N=h, C=c, elems=x.numel()), call, (x, out, scale, shift, ws)))
    return out


def lpf_down2(x: torch.Tensor, gn_stats: bool = False) -> torch.Tensor:
    """LPF_RFFT(0.5)(x)[..., ::2, ::2] (af_blocks.py:149-150) on NHWC x.  ``gn_stats``: also emit the GroupNorm
    partial sums of the output (attached to the returned tensor, as ``conv2d(gn_stats=True)`` does)."""
    _chk(x, "x")
    b, h2, w2, c = x.shape
    if h2 % 2 or w2 % 2:
        raise _lib.AfldmError("lpf_down2: even input size expected")
    out = torch.empty((b, h2 // 2, w2 // 2, c), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    if gn_stats and h2 // 2 <= 16 and h2 == w2:
        gn = torch.empty((b, 1, c, 2), dtype=torch.float32, device=x.device)

        def call_gn():
            return L.afldm_lpf_down2_gn_f32(x.data_ptr(), out.data_ptr(), b, h2 // 2, w2 // 2, c, gn.data_ptr(), _stream())

        code = call_gn()
        if code not in _UNSUPPORTED:
            _lib.check(code, "lpf_down2_gn")
            if _recorder is not None:
                _recorder.append(("lpf_down2", dict(B=b, N=h2 // 2, C=c, elems=x.numel(), gn=1), call_gn, (x, out, gn)))
            out._afldm_gn = (gn, 1, c)
            return out
    need = L.afldm_resample_workspace_floats(2, b, h2 // 2, w2 // 2, c)
    ws = scratch(x.device, need) if need else None

    def call():
        return L.afldm_lpf_down2_f32(x.data_ptr(), out.data_ptr(), b, h2 // 2, w2 // 2, c, _ptr(ws), need, _stream())

    code = call()
    if code in _UNSUPPORTED and GENERAL_PLANES and h2 == w2:
        return _general_resample(x, "down2")
    _lib.check(code, "lpf_down2")
    if _recorder is not None:
        _recorder.append(("lpf_down2", dict(B=b, N=h2 // 2, C=c, elems=x.numel()), call, (x, out, ws)))
    return out


# ------------------------------------------------------------------------- norm / activation
def groupnorm_affine(x: torch.Tensor, groups: int, eps: float, gamma: Optional[torch.Tensor],
                     beta: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """GroupNorm statistics of NHWC x folded into per-(b,c) scale / shift, each [B, C]."""
    _chk(x, "x")
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    L = _lib.lib()
    one, two = getattr(x, "_afldm_gn", None), getattr(x, "_afldm_gn2", None)
    if one is not None or two is not None:
        # the kernel(s) that produced x already emitted its partial sums: no pass over x
        (pa, sa, ca), (pb, sb, cb) = (one, (None, 0, 0)) if one is not None else two
        if ca + cb == c:
            ss = torch.empty((2, b, c), dtype=torch.float32, device=x.device)
            _run("groupnorm_finalize", dict(B=b, HW=hw, C=c, elems=b * c),
                 lambda: L.afldm_groupnorm_finalize_f32(pa.data_ptr(), sa, ca, _ptr(pb), sb, cb, b, hw, groups,
                                                        float(eps), _ptr(gamma), _ptr(beta), ss[0].data_ptr(),
                                                        ss[1].data_ptr(), _stream()), (pa, pb, gamma, beta, ss))
            return ss[0], ss[1]
    need = L.afldm_groupnorm_scratch_floats(b, hw, c)
    part = scratch(x.device, need) if need else None
    ss = torch.empty((2, b, c), dtype=torch.float32, device=x.device)
    _run("groupnorm_affine", dict(B=b, HW=hw, C=c, elems=x.numel()),
         lambda: L.afldm_groupnorm_affine_f32(x.data_ptr(), b, hw, c, groups, float(eps), _ptr(gamma), _ptr(beta),
                                              ss[0].data_ptr(), ss[1].data_ptr(), _ptr(part), _stream()),
         (x, gamma, beta, ss, part))
    return ss[0], ss[1]


def affine_act(x: torch.Tensor, scale: Optional[torch.Tensor], shift: Optional[torch.Tensor],
               act: str = "silu", out: Optional[torch.Tensor] = None, out_half: bool = False) -> torch.Tensor:
    _chk(x, "x")
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    L = _lib.lib()
    if out_half and out is None:
        outh = torch.empty(x.shape, dtype=torch.float16, device=x.device)
        _run("affine_act", dict(elems=x.numel(), f16out=1),
             lambda: L.afldm_affine_act_f16out(x.data_ptr(), outh.data_ptr(), b, hw, c, ACT[act], _ptr(scale), _ptr(shift),
                                               _stream()), (x, outh, scale, shift))
        return outh
    if out is None:
        out = torch.empty_like(x)
    _run("affine_act", dict(elems=x.numel()),
         lambda: L.afldm_affine_act_f32(x.data_ptr(), out.data_ptr(), b, hw, c, ACT[act], _ptr(scale), _ptr(shift),
                                        _stream()), (x, out, scale, shift))
    return out


def groupnorm_act(x: torch.Tensor, groups: int, eps: float, gamma: Optional[torch.Tensor],
                  beta: Optional[torch.Tensor], act: str = "identity", out_half: bool = False) -> torch.Tensor:
    """act(GroupNorm(x)) materialised (the normalised input of an attention block).  One launch when the producer
    of x emitted its GroupNorm partial sums (``FUSE_GN_PROLOGUE``), else statistics / finalize + ``affine_act``."""
    _chk(x, "x")
    b, c = x.shape[0], x.shape[-1]
    hw = x.numel() // (b * c)
    one, two = getattr(x, "_afldm_gn", None), getattr(x, "_afldm_gn2", None)
    if FUSE_GN_PROLOGUE and (one is not None or two is not None) and c % 4 == 0 and 2 * c + 2 * groups <= 12288:
        (pa, sa, ca), (pb, sb, cb) = (one, (None, 0, 0)) if one is not None else two
        if ca + cb == c:
            L = _lib.lib()
            if out_half:        # fp16 result: the only consumer is a tensor-core projection
                outh = torch.empty(x.shape, dtype=torch.float16, device=x.device)
                _run("affine_act", dict(elems=x.numel(), fused_gn=1, f16out=1),
                     lambda: L.afldm_affine_act_gn_f16out(x.data_ptr(), outh.data_ptr(), b, hw, c, ACT[act], pa.data_ptr(),
                                                          sa, ca, _ptr(pb), sb, cb, groups, float(eps), _ptr(gamma),
                                                          _ptr(beta), _stream()), (x, outh, pa, pb, gamma, beta))
                return outh
            out = torch.empty_like(x)
            _run("affine_act", dict(elems=x.numel(), fused_gn=1),
                 lambda: L.afldm_affine_act_gn_f32(x.data_ptr(), out.data_ptr(), b, hw, c, ACT[act], pa.data_ptr(), sa, ca,
                                                   _ptr(pb), sb, cb, groups, float(eps), _ptr(gamma), _ptr(beta),
                                                   _stream()), (x, out, pa, pb, gamma, beta))
            return out
    scale, shift = groupnorm_affine(x, groups, eps, gamma, beta)
    return affine_act(x, scale, shift, act=act, out_half=out_half and c % 4 == 0)


# ------------------------------------------------------------------------- conv / linear
def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d weight [Cout,Cin,kh,kw] (or nn.Linear [Cout,Cin]) -> packed [Cout][kh*kw][Cin]."""
    if w.ndim == 2:
        return w.detach().contiguous()
    co, ci, kh, kw = w.shape
    return w.detach().permute(0, 2, 3, 1).reshape(co, kh * kw, ci).contiguous()


def conv2d(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], ksize: int,
           row_add: Optional[torch.Tensor] = None, residual: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None, algo: Optional[str] = None, gn_stats: bool = False) -> torch.Tensor:
    """Stride-1 'same' convolution (k = 1 or 3) of NHWC x with the fused epilogue
    ``+ bias + row_add[b] + residual``.  ``gn_stats``: also emit the GroupNorm partial sums of the output
    (tensor-core path), attached to the returned tensor for ``groupnorm_affine`` to pick up.  ``x``, ``residual`` and ``out`` may be channel slices of
    wider NHWC buffers (last-dim stride 1, pixel pitch = stride(-2))."""
    L = _lib.lib()
    if row_add is not None:
        join_deferred()
    b, h, w_, cin = x.shape
    cout = w_packed.shape[0]
    if x.stride(-1) != 1 or not x.is_cuda or x.dtype not in (torch.float32, torch.float16):
        raise _lib.AfldmError("conv2d: fp32 (or fp16) CUDA NHWC tensor expected")
    f16in = x.dtype == torch.float16
    if f16in != (w_packed.dtype == torch.float16):
        raise _lib.AfldmError("conv2d: an fp16 activation needs the fp16-packed weight (packing.conv_params_f16) and vice versa")
    x_pitch = _pitch(x)
    if out is None:
        out = torch.empty((b, h, w_, cout), dtype=torch.float32, device=x.device)
    y_pitch = _pitch(out)
    res_pitch = _pitch(residual) if residual is not None else 0
    ra_pitch = 0
    if row_add is not None:
        if row_add.ndim != 2 or row_add.shape != (b, cout) or row_add.stride(1) != 1:
            raise _lib.AfldmError("conv2d: row_add must be [B, Cout] with unit column stride")
        ra_pitch = row_add.stride(0) if b > 1 else cout
    name = algo or _default_conv_algo
    meta = dict(B=b, H=h, W=w_, Cin=cin, Cout=cout, k=ksize, flops=2.0 * b * h * w_ * cout * ksize * ksize * cin)
    keep = (x, w_packed, bias, row_add, residual, out)

    for stale in ("_afldm_gn", "_afldm_gn2"):
        if hasattr(out, stale):
            delattr(out, stale)

    if f16in:
        # fp16 operands (tcgen05.mma.kind::f16): tensor-core path only, no SIMT form
        need = L.afldm_conv2d_workspace_floats(b, h, w_, cin, cout, ksize, 2)
        ws = scratch(x.device, need) if need else None
        slots = L.afldm_conv2d_gn_slots(b, h, w_, cin, cout, ksize, 2) if gn_stats else 0
        gn = torch.empty((b, slots, cout, 2), dtype=torch.float32, device=x.device) if slots else None
        meta["f16in"] = 1
        _run("conv2d_f16", meta,
             lambda: L.afldm_conv2d_f16in_f32(x.data_ptr(), x_pitch, w_packed.data_ptr(), _ptr(bias), _ptr(row_add), ra_pitch,
                                              _ptr(residual), res_pitch, out.data_ptr(), y_pitch, b, h, w_, cin, cout, ksize,
                                              _ptr(ws), need, _ptr(gn), _stream()), keep + (ws, gn))
        if gn is not None:
            out._afldm_gn = (gn, slots, cout)
        return out

    def call(a: int, ws, need: int, gn=None):
        return L.afldm_conv2d_f32(x.data_ptr(), x_pitch, w_packed.data_ptr(), _ptr(bias), _ptr(row_add), ra_pitch,
                                  _ptr(residual), res_pitch, out.data_ptr(), y_pitch, b, h, w_, cin, cout, ksize,
                                  a, _ptr(ws), need, _ptr(gn), _stream())

    if CONV_ALGO[name] == 1:
        need = L.afldm_conv2d_workspace_floats(b, h, w_, cin, cout, ksize, 1)
        ws = scratch(x.device, need) if need else None
        slots = L.afldm_conv2d_gn_slots(b, h, w_, cin, cout, ksize, 1) if gn_stats else 0
        gn = torch.empty((b, slots, cout, 2), dtype=torch.float32, device=x.device) if slots else None
        code = call(1, ws, need, gn)
        if code != -3:      # AFLDM_E_NOKERNEL: shape outside the tensor-core family -> exact SIMT kernel
            _lib.check(code, "conv2d[tf32]")
            if _recorder is not None:
                _recorder.append(("conv2d_tf32", meta, lambda: call(1, ws, need, gn), keep + (ws, gn)))
            if gn is not None:
                out._afldm_gn = (gn, slots, cout)
            return out
    need0 = L.afldm_conv2d_workspace_floats(b, h, w_, cin, cout, ksize, 0)
    ws0 = scratch(x.device, need0) if need0 else None
    _run("conv2d_simt", meta, lambda: call(0, ws0, need0), keep + (ws0,))
    return out


F16_ATTENTION = os.environ.get("AFLDM_ATTN_F16", "1") == "1"
# TF32 class: the filtered activation in front of a resnet convolution stores fp16 (the 11 significant bits the tensor
# core would keep of it anyway) and the convolution runs tcgen05.mma.kind::f16 on fp16-packed weights
F16_CONV = os.environ.get("AFLDM_CONV_F16", "1") == "1"


def conv2d_f16out(x: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], ksize: int) -> Optional[torch.Tensor]:
    """fp16(conv(x) + bias) on the tcgen05 path (the q | k | v projection in front of ``attention_f16``).
    Returns None when this shape has no fp16 epilogue (split-K layers, shapes outside the tensor-core family):
    the caller then uses ``conv2d`` + ``attention``."""
    if not x.is_cuda or x.dtype not in (torch.float32, torch.float16) or w_packed.dtype != x.dtype or \
            not w_packed.is_contiguous():
        raise _lib.AfldmError("conv2d_f16out: CUDA input and packed weight of one dtype (fp32 or fp16) expected")
    f16in = x.dtype == torch.float16
    b, h, w_, cin = x.shape
    cout = w_packed.shape[0]
    x_pitch = _pitch(x)
    out = torch.empty((b, h, w_, cout), dtype=torch.float16, device=x.device)
    L = _lib.lib()

    def call():
        fn = L.afldm_conv2d_f16in_f16out if f16in else L.afldm_conv2d_f16out
        return fn(x.data_ptr(), x_pitch, w_packed.data_ptr(), _ptr(bias), out.data_ptr(), cout,
                  b, h, w_, cin, cout, ksize, _stream())

    code = call()
    if code == -3:
        return None
    _lib.check(code, "conv2d_f16out")
    if _recorder is not None:
        _recorder.append(("conv2d_f16" if f16in else "conv2d_tf32",
                          dict(B=b, H=h, W=w_, Cin=cin, Cout=cout, k=ksize, f16out=1,
                                              flops=2.0 * b * h * w_ * cout * cin * ksize * ksize), call,
                          (x, w_packed, bias, out)))
    return out


def attention_f16(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int,
                  out: Optional[torch.Tensor] = None, out_half: bool = False) -> torch.Tensor:
    """``attention`` for fp16 q / k / v (column slices of the fp16 fused-QKV buffer); fp32 output, or fp16
    (``out_half``) when a tensor-core projection is the only consumer."""
    b, nq, cd = q.shape
    bkv, nk, _ = k.shape
    d = cd // heads
    for t in (q, k, v):
        if t.dtype != torch.float16 or not t.is_cuda:
            raise _lib.AfldmError("attention_f16: fp16 CUDA tensors expected")
        if t.stride(-1) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
            raise _lib.AfldmError("attention_f16: rows must be densely pitched")
    if k.stride(1) != v.stride(1):
        raise _lib.AfldmError("attention_f16: k and v must share a pitch")
    if out is None:
        out = torch.empty((b, nq, cd), dtype=torch.float16 if out_half else torch.float32, device=q.device)
    L = _lib.lib()
    fn = L.afldm_attention_f16_f16out if out.dtype == torch.float16 else L.afldm_attention_f16
    _run("attention_f16", dict(B=b, Nq=nq, Nk=nk, heads=heads, d=d, flops=4.0 * b * heads * nq * nk * d),
         lambda: fn(q.data_ptr(), q.stride(1), k.data_ptr(), v.data_ptr(), k.stride(1),
                    out.data_ptr(), out.stride(1), b, bkv, nq, nk, heads, d, _stream()),
         (q, k, v, out))
    return out


def _pitch(t: torch.Tensor) -> int:
    """Pixel pitch (floats) of an NHWC tensor or of a channel slice of a wider NHWC buffer."""
    b, h, w, c = t.shape
    if t.stride(3) != 1 and c > 1:
        raise _lib.AfldmError(f"NHWC tensor with unit channel stride expected, got strides {t.stride()}")
    p = t.stride(2) if w > 1 else (t.stride(1) if h > 1 else (t.stride(0) if b > 1 else c))
    ok = p >= c and (w == 1 or t.stride(2) == p) and (h == 1 or t.stride(1) == w * p) and \
        (b == 1 or t.stride(0) == h * w * p)
    if not ok:
        raise _lib.AfldmError(f"unsupported NHWC strides {t.stride()} for shape {tuple(t.shape)}")
    return p


def linear_rows(x: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], act_in: str = "identity",
                act_out: str = "identity") -> torch.Tensor:
    """y = act_out(act_in(x) @ w.T + bias) for a few rows (time-embedding MLP)."""
    if _deferred_join is not None and torch.cuda.current_stream(x.device) == _deferred_join[0]:
        join_deferred()          # a resnet projecting the embedding itself, on the main stream
    _chk(x, "x")
    _chk(w, "w")
    m, k = x.shape
    n = w.shape[0]
    y = torch.empty((m, n), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    rows = max(1, min(64, (200 * 1024) // (4 * k)))       # the kernel keeps its rows of x in shared memory: <= 64 at a time
    for m0 in range(0, m, rows):
        xs, ys = x[m0:m0 + rows], y[m0:m0 + rows]
        mm = xs.shape[0]
        _run("linear_rows", dict(M=mm, K=k, N=n, bytes=4.0 * n * k),
             lambda xs=xs, ys=ys, mm=mm: L.afldm_linear_rows_f32(xs.data_ptr(), w.data_ptr(), _ptr(bias), ys.data_ptr(), mm, k, n,
                                                                 ACT[act_in], ACT[act_out], _stream()), (xs, w, bias, ys))
    return y


# ------------------------------------------------------------------------- attention
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int,
              out: Optional[torch.Tensor] = None, algo: Optional[str] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(d)) v per head.  q [B,Nq,heads*d], k/v [Bkv,Nk,heads*d]; each may be a
    column slice of a wider (e.g. fused QKV) buffer.  B % Bkv == 0: batch b uses K/V batch b // (B/Bkv)."""
    b, nq, cd = q.shape
    bkv, nk, _ = k.shape
    d = cd // heads
    for t in (q, k, v):
        if t.stride(-1) != 1 or t.stride(0) != t.shape[1] * t.stride(1):
            raise _lib.AfldmError("attention: rows must be densely pitched")
    if k.stride(1) != v.stride(1):
        raise _lib.AfldmError("attention: k and v must share a pitch")
    if out is None:
        out = torch.empty((b, nq, cd), dtype=torch.float32, device=q.device)
    L = _lib.lib()
    # precision class follows the convolution path: 'tf32' -> tensor-core products, 'simt' -> exact fp32
    a = CONV_ALGO[algo or _default_conv_algo] if d % 8 == 0 else 0
    _run("attention_tf32" if a else "attention", dict(B=b, Nq=nq, Nk=nk, heads=heads, d=d, flops=4.0 * b * heads * nq * nk * d),
         lambda: L.afldm_attention_f32(q.data_ptr(), q.stride(1), k.data_ptr(), v.data_ptr(), k.stride(1),
                                       out.data_ptr(), out.stride(1), b, bkv, nq, nk, heads, d, a, _stream()),
         (q, k, v, out))
    return out


def softmax_rows_(x: torch.Tensor, scale: float) -> torch.Tensor:
    """In-place softmax(scale * x) over the last dim of a contiguous matrix."""
    _chk(x, "x")
    cols = x.shape[-1]
    L = _lib.lib()
    _run("softmax_rows", dict(elems=x.numel()),
         lambda: L.afldm_softmax_rows_f32(x.data_ptr(), x.numel() // cols, cols, cols, float(scale), _stream()), (x,))
    return x


def attention_gemm(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int) -> torch.Tensor:
    """Attention for head dims the register-resident kernel does not cover (VAE mid block: one head
    of 512): per (batch, head)  S = q k^T  ->  row softmax  ->  O = S v, the two products running
    through the implicit-GEMM kernel as 1x1 convolutions."""
    b, nq, cd = q.shape
    bkv, nk, _ = k.shape
    d = cd // heads
    rep = b // bkv
    out = torch.empty((b, nq, cd), dtype=torch.float32, device=q.device)
    for bi in range(b):
        kb = bi // rep
        for hd in range(heads):
            sl = slice(hd * d, (hd + 1) * d)
            qh = q[bi:bi + 1, :, None, sl]                      # [1, Nq, 1, d]  (x of the "conv")
            kh = k[kb, :, sl].contiguous()                      # [Nk, d]        (weights [Cout=Nk][Cin=d])
            s = conv2d(qh, kh, None, 1)                         # [1, Nq, 1, Nk]
            softmax_rows_(s.view(nq, nk), d ** -0.5)
            vt = to_nchw_contiguous(v[kb:kb + 1, :, None, sl].contiguous()).view(d, nk)   # v^T [d, Nk]
            conv2d(s, vt, None, 1, out=out[bi:bi + 1, :, None, sl])
    return out


# ------------------------------------------------------------------------- step ops
def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    _chk(t, "t")
    out = torch.empty((t.shape[0], dim), dtype=torch.float32, device=t.device)
    L = _lib.lib()
    _run("timestep_embedding", dict(elems=out.numel()),
         lambda: L.afldm_timestep_embedding_f32(t.data_ptr(), out.data_ptr(), t.shape[0], dim, _stream()), (t, out))
    return out


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _chk(a, "a")
    _chk(b, "b")
    ca, cb = a.shape[-1], b.shape[-1]
    pixels = a.numel() // ca
    y = torch.empty(a.shape[:-1] + (ca + cb,), dtype=torch.float32, device=a.device)
    L = _lib.lib()
    _run("concat_channels", dict(elems=y.numel()),
         lambda: L.afldm_concat_channels_f32(a.data_ptr(), ca, b.data_ptr(), cb, y.data_ptr(), pixels, _stream()),
         (a, b, y))
    ga, gb = getattr(a, "_afldm_gn", None), getattr(b, "_afldm_gn", None)
    if ga is not None and gb is not None:
        y._afldm_gn2 = (ga, gb)         # statistics of a concat = the statistics of its two halves
    return y


def pad_channels(x: torch.Tensor, cpad: int) -> torch.Tensor:
    """NHWC x [B,H,W,C] -> [B,H,W,cpad] with zeros in the new channels."""
    _chk(x, "x")
    c = x.shape[-1]
    y = torch.empty(x.shape[:-1] + (cpad,), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    _run("pad_channels", dict(elems=y.numel()),
         lambda: L.afldm_pad_channels_f32(x.data_ptr(), c, y.data_ptr(), cpad, x.numel() // c, _stream()), (x, y))
    return y


def axpby(x: torch.Tensor, e: torch.Tensor, cx, ce, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = cx * x + ce * e.  cx/ce floats, or a device tensor ``coef`` of 2 floats passed as cx (ce None)."""
    _chk(x, "x")
    _chk(e, "e")
    if out is None:
        out = torch.empty_like(x)
    L = _lib.lib()
    n = x.numel()
    if isinstance(cx, torch.Tensor):
        _run("axpby", dict(elems=n),
             lambda: L.afldm_axpby_dev_f32(x.data_ptr(), e.data_ptr(), out.data_ptr(), cx.data_ptr(), n, _stream()),
             (x, e, out, cx))
    else:
        fx, fe = float(cx), float(ce)
        _run("axpby", dict(elems=n),
             lambda: L.afldm_axpby_f32(x.data_ptr(), e.data_ptr(), out.data_ptr(), fx, fe, n, _stream()), (x, e, out))
    return out


def geglu(proj: torch.Tensor) -> torch.Tensor:
    """diffusers GEGLU on the projection [..., 2H] -> [..., H]: ``proj[..., :H] * gelu(proj[..., H:])`` (exact GELU)."""
    _chk(proj, "proj")
    h = proj.shape[-1] // 2
    rows = proj.numel() // (2 * h)
    y = torch.empty(proj.shape[:-1] + (h,), dtype=torch.float32, device=proj.device)
    L = _lib.lib()
    _run("geglu", dict(elems=y.numel()), lambda: L.afldm_geglu_f32(proj.data_ptr(), y.data_ptr(), rows, h, _stream()), (proj, y))
    return y


def layer_norm(x: torch.Tensor, weight: Optional[torch.Tensor], bias: Optional[torch.Tensor], eps: float) -> torch.Tensor:
    """nn.LayerNorm over the last dim of [..., C] (SD-1.5 transformer blocks): every token is one "image" of one pixel for
    the GroupNorm statistics kernel (groups = 1), the affine is the same per-channel gamma / beta."""
    _chk(x, "x")
    c = x.shape[-1]
    tokens = x.numel() // c
    y = groupnorm_act(x.view(tokens, 1, 1, c), 1, eps, weight, bias, act="identity")
    return y.view(x.shape)


def slot_copy(table: torch.Tensor, buf: torch.Tensor, slot: torch.Tensor, store: bool) -> None:
    """``table[slot] = buf`` (store) or ``buf = table[slot]`` with ``slot`` a DEVICE int32 scalar - graph-capturable
    per-timestep storage of the cross-frame attention maps (cross_frame_attn.py:78-97)."""
    _chk(table, "table")
    _chk(buf, "buf")
    if slot.dtype != torch.int32 or not slot.is_cuda or table.ndim < 2 or table[0].numel() != buf.numel():
        raise _lib.AfldmError("slot_copy: table [slots, ...] / buf of one slot / device int32 slot expected")
    n = buf.numel()
    L = _lib.lib()
    _run("slot_copy", dict(elems=n),
         lambda: L.afldm_slot_copy_f32(table.data_ptr(), buf.data_ptr(), n, slot.data_ptr(), int(bool(store)), _stream()),
         (table, buf, slot))


def upfirdn2d(x: torch.Tensor, f: torch.Tensor, up=1, down=1, padding=(0, 0, 0, 0), flip_filter=False,
              gain=1.0) -> torch.Tensor:
    """StyleGAN3 upfirdn2d on NCHW x (upfirdn2d.py:118-162); padding = (x0, x1, y0, y1); f 1-D or 2-D."""
    _chk(x, "x")
    if f.ndim == 1:
        f = torch.outer(f, f)
    f = _chk(f.contiguous(), "f")
    b, c, h, w = x.shape
    fh, fw = f.shape
    px0, px1, py0, py1 = padding
    ow = (w * up + px0 + px1 - fw + down) // down
    oh = (h * up + py0 + py1 - fh + down) // down
    y = torch.empty((b, c, oh, ow), dtype=torch.float32, device=x.device)
    L = _lib.lib()
    _run("upfirdn2d", dict(elems=y.numel()),
         lambda: L.afldm_upfirdn2d_f32(x.data_ptr(), f.data_ptr(), y.data_ptr(), b, c, h, w, fh, fw, up, up, down, down,
                                       px0, px1, py0, py1, int(bool(flip_filter)), float(gain), _stream()), (x, f, y))
    return y
