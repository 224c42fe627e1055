"""afldm/shift_utils/shifters.py of the reference (:31-206)."""
from afldm_b200.shift_utils.shifters import (FILTER_CHOICES, ImageShifter, fourier_shift_batch,  # noqa: F401
                                             gen_random_offset, gen_valid_mask)
