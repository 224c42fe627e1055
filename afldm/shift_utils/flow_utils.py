"""afldm/shift_utils/flow_utils.py of the reference: the part the shift measurement uses (:34-86)."""
from afldm_b200.shift_utils.flow_utils import bilinear_sample, coords_grid, flow_warp  # noqa: F401
