from .unet_2d import UNet2DModel  # noqa: F401
from .af_vae import AliasFreeAutoencoderKL, AutoencoderKL  # noqa: F401
from .unet_2d_condition import UNet2DConditionModel  # noqa: F401
