"""afldm/schedulers/i2sb_scheduler.py of the reference (:142-531)."""
from afldm_b200.schedulers.i2sb import I2SBScheduler  # noqa: F401
