from afldm_b200.models.blocks import Attention, AttnProcessor2_0  # noqa: F401
