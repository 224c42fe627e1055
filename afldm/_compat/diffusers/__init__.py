"""Minimal stand-in for the ``diffusers`` names the reference's entry scripts import directly
(scripts/shift_ldm_ffhq.py:5, scripts/video_editing.py:4).  Only on ``sys.path`` behind the real package:
it is found when ``diffusers`` itself is not installed (this image)."""
from . import models, utils  # noqa: F401

__version__ = "0.32.1+afldm_b200.compat"
