"""afldm/shift_utils/metrics.py of the reference (:5-19)."""
from afldm_b200.shift_utils.metrics import mask_mse, mask_psnr, psnr  # noqa: F401
