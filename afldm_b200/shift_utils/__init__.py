from .shifters import (FILTER_CHOICES, ImageShifter, fourier_shift_batch, gen_random_offset, gen_valid_mask,  # noqa: F401
                       shift_matrix)
from .metrics import mask_mse, mask_psnr, psnr  # noqa: F401
