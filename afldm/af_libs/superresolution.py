"""Reference module path ``afldm.af_libs.superresolution`` -> ``afldm_b200.af_libs.superresolution``."""
from afldm_b200.af_libs.superresolution import *            # noqa: F401,F403
from afldm_b200.af_libs.superresolution import build_sr4x, build_sr_bicubic, build_sr_pool   # noqa: F401
