"""ctypes binding of ``libafldm_b200.so`` - the C-ABI declared in ``include/afldm_b200.h``.

The library is the product: there is no Python / PyTorch fallback.  If it has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``make -C afldm_b200/csrc``) every
op raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libafldm_b200.so")
CSRC_DIR = os.path.join(_HERE, "csrc")

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_ll = C.c_longlong
_sz = C.c_size_t

# name -> (restype, argtypes); mirrors include/afldm_b200.h one to one.
SIGNATURES = {
    "afldm_abi_version": (_i, []),
    "afldm_launch_count": (C.c_ulonglong, []),
    "afldm_error_string": (C.c_char_p, [_i]),
    "afldm_resample_workspace_floats": (_sz, [_i, _i, _i, _i, _i]),
    "afldm_filtered_act_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "afldm_filtered_act_gn_f32": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p, _i, _i, _i, _f, _p, _p, _p]),
    "afldm_filtered_act_gn_cat_f32": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _i, _i, _f, _p, _p, _p]),
    "afldm_filtered_act_gn_f16out": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _i, _i, _p, _i, _i, _i, _f, _p, _p, _p]),
    "afldm_filtered_act_gn_cat_f16out": (_i, [_p, _p, _p, _i, _i, _i, _i, _i, _i, _p, _i, _p, _i, _i, _f, _p, _p, _p]),
    "afldm_up2_ideal_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "afldm_up2_ideal_f16out": (_i, [_p, _p, _i, _i, _i, _i, _p, _sz, _p]),
    "afldm_filtered_act_f16out": (_i, [_p, _p, _i, _i, _i, _i, _i, _p, _p, _p, _sz, _p]),
    "afldm_filtered_act_tc": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _p, _p, _p]),
    "afldm_affine_act_f16out": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "afldm_lpf_down2_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _sz, _p]),
    "afldm_lpf_down2_gn_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p]),
    "afldm_groupnorm_scratch_floats": (_sz, [_i, _i, _i]),
    "afldm_groupnorm_affine_f32": (_i, [_p, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p, _p]),
    "afldm_affine_act_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _p, _p]),
    "afldm_affine_act_gn_f32": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _i, _p, _i, _i, _i, _f, _p, _p, _p]),
    "afldm_affine_act_gn_f16out": (_i, [_p, _p, _i, _i, _i, _i, _p, _i, _i, _p, _i, _i, _i, _f, _p, _p, _p]),
    "afldm_conv2d_workspace_floats": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "afldm_conv2d_supported": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "afldm_conv2d_plan": (_i, [_i, _i, _i, _i, _i, _i, _i, C.POINTER(C.c_int)]),
    "afldm_conv2d_gn_slots": (_i, [_i, _i, _i, _i, _i, _i, _i]),
    "afldm_conv2d_f32": (_i, [_p, _i, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p, _p]),
    "afldm_conv2d_f16in_f32": (_i, [_p, _i, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p, _sz, _p, _p]),
    "afldm_conv2d_cat_f32": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _i, _p, _i, _p, _i, _i, _i, _i, _i, _i, _p, _sz, _p, _p]),
    "afldm_groupnorm_finalize_f32": (_i, [_p, _i, _i, _p, _i, _i, _i, _i, _i, _f, _p, _p, _p, _p, _p]),
    "afldm_linear_rows_f32": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "afldm_attention_f32": (_i, [_p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "afldm_attention_f16": (_i, [_p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "afldm_attention_f16_f16out": (_i, [_p, _i, _p, _p, _i, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "afldm_conv2d_f16in_f16out": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "afldm_conv2d_f16out": (_i, [_p, _i, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _p]),
    "afldm_softmax_rows_f32": (_i, [_p, _ll, _i, _i, _f, _p]),
    "afldm_timestep_embedding_f32": (_i, [_p, _p, _i, _i, _p]),
    "afldm_concat_channels_f32": (_i, [_p, _i, _p, _i, _p, _ll, _p]),
    "afldm_pad_channels_f32": (_i, [_p, _i, _p, _i, _ll, _p]),
    "afldm_nchw_to_nhwc_f32": (_i, [_p, _p, _i, _i, _i, _p]),
    "afldm_nhwc_to_nchw_f32": (_i, [_p, _p, _i, _i, _i, _p]),
    "afldm_axpby_f32": (_i, [_p, _p, _p, _f, _f, _ll, _p]),
    "afldm_axpby_dev_f32": (_i, [_p, _p, _p, _p, _ll, _p]),
    "afldm_slot_copy_f32": (_i, [_p, _p, _ll, _p, _i, _p]),
    "afldm_geglu_f32": (_i, [_p, _p, _ll, _i, _p]),
    "afldm_act_bwd_mul_f32": (_i, [_p, _p, _p, _ll, _i, _p]),
    "afldm_plane_sep_transform_workspace_floats": (_sz, [_i, _i, _i]),
    "afldm_plane_sep_transform_f32": (_i, [_p, _p, _p, _p, _p, _sz, _i, _i, _i, _i, _i, _i, _p]),
    "afldm_upfirdn2d_f32": (_i, [_p, _p, _p] + [_i] * 15 + [_f, _p]),
}

_lib = None


class AfldmError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile the CUDA sources for sm_100a into ``afldm_b200/lib/libafldm_b200.so``."""
    r = subprocess.run(["make", "-C", CSRC_DIR, "-j8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise AfldmError("building libafldm_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library (loads on first use; fails loudly when it is missing)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise AfldmError(
                f"{LIB_PATH} is missing: build it with `make -C afldm_b200/csrc` "
                "(afldm_b200 has no CPU / PyTorch fallback)")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        if handle.afldm_abi_version() != 1:
            raise AfldmError("libafldm_b200.so: unexpected ABI version")
        _lib = handle
    return _lib


_SYNC_EACH = os.environ.get("AFLDM_SYNC_EACH", "0") == "1"      # debugging aid: wait for every launch (finds the kernel that hangs)


def check(code: int, what: str = "") -> None:
    if _SYNC_EACH and code == 0:
        import torch
        if not torch.cuda.is_current_stream_capturing():
            torch.cuda.synchronize()
    if code != 0:
        msg = lib().afldm_error_string(code).decode()
        raise AfldmError(f"{what or 'afldm call'} failed ({code}): {msg}")


def launch_count() -> int:
    return int(lib().afldm_launch_count())
