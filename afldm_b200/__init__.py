"""afldm_b200 - B200-native (sm_100a) implementation of the AF-LDM denoising hot path.

Host-side mirror of the reference's module surface (``afldm.af_modules``, ``afldm.af_libs.ideal_lpf``,
``afldm.models``, ``afldm.pipelines``) on top of the C-ABI library ``libafldm_b200.so``
(``include/afldm_b200.h``).  No CPU / PyTorch fallback: ops raise when the library or a GPU is missing.
"""
from . import _lib  # noqa: F401

__all__ = ["_lib", "ops"]
