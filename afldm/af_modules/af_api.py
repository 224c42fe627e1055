"""afldm/af_modules/af_api.py of the reference (:9-93), sm_100a kernels underneath."""
from afldm_b200.af_modules.af_api import (make_af_controlnet, make_af_unet, make_af_vae,  # noqa: F401
                                          make_af_vae_from_config, replace_downsampler, replace_upsampler,
                                          wrap_nonlinearity, wrap_resblock_nonlinearity)
