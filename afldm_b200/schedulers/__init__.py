from .ddim import DDIMScheduler  # noqa: F401
from .i2sb import I2SBScheduler  # noqa: F401
