from .cross_frame_attn import AttnState, CrossFrameAttnProcessor, get_unet_attn_processors, set_unet_attn_processor  # noqa: F401
from .ldm_pipeline import MyLDMPipeline  # noqa: F401
from .i2sb_pipeline import I2SBLDMPipeline  # noqa: F401
from .video_equiv_editing_pipeline import VideoEquivariantEditingPipeline  # noqa: F401
