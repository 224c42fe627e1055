"""afldm/af_libs/torch_utils/ops/upfirdn2d.py of the reference (:70-387), sm_100a kernel underneath."""
from afldm_b200.af_libs.upfirdn2d import (downsample2d, filter2d, setup_filter, upfirdn2d,  # noqa: F401
                                          upsample2d)
