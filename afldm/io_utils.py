"""afldm/io_utils.py of the reference (:10-61)."""
from afldm_b200.io_utils import image_to_tensor, save_gif_from_tensors  # noqa: F401
