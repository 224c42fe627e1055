"""``afldm`` - the reference's import surface (SingleZombie/AFLDM module paths), backed by ``afldm_b200``.

``PYTHONPATH=/root/repo:/root/repo/afldm/_compat python <reference>/scripts/shift_ldm_ffhq.py`` resolves every
``afldm.*`` import of the reference's entry scripts against this package: the module-swap API
(``afldm.af_modules.af_api``), the alias-free blocks, ``afldm.af_libs.ideal_lpf``, ``afldm.models.af_vae``,
``afldm.pipelines.{ldm_pipeline,cross_frame_attn,i2sb_pipeline}``, ``afldm.schedulers.i2sb_scheduler``,
``afldm.shift_utils.{shifters,metrics,flow_utils}`` and ``afldm.io_utils`` - same names, signatures and error
behaviour, sm_100a kernels underneath (no CPU fallback).  ``afldm/_compat/diffusers`` is a minimal stand-in for the
few ``diffusers`` names the scripts import themselves; it is used only when the real package is not installed.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

import afldm_b200 as _impl  # noqa: F401  (fails loudly when the product package is missing)

if _ilu.find_spec("diffusers") is None:          # real diffusers absent: expose the stand-in under its name
    _compat = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "_compat")
    if _compat not in _sys.path:
        _sys.path.append(_compat)
