"""afldm/af_libs/ideal_lpf.py of the reference (:12-172), sm_100a kernels underneath."""
from afldm_b200.af_libs.ideal_lpf import (LPF_RECON_RFFT, LPF_RFFT, UpsampleRFFT, create_lpf_rect,  # noqa: F401
                                          create_recon_rect, subpixel_shift)
