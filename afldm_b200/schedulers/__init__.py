from .ddim import DDIMScheduler  # noqa: F401
