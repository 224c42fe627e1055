"""afldm/pipelines/cross_frame_attn.py of the reference (:6-190)."""
from afldm_b200.pipelines.cross_frame_attn import (AttnState, CrossFrameAttnProcessor,  # noqa: F401
                                                   get_unet_attn_processors, set_unet_attn_processor)
