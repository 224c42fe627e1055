"""afldm/af_modules/af_blocks.py of the reference (:12-152), sm_100a kernels underneath."""
from afldm_b200.af_modules.af_blocks import (AliasFreeDownsample2D, AliasFreeUpsample2D,  # noqa: F401
                                             WarpedNonlinearity)
