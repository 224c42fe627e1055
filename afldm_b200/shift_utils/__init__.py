from .shifters import ImageShifter, gen_valid_mask, shift_matrix  # noqa: F401
