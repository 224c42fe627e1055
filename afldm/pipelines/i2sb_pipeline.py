"""afldm/pipelines/i2sb_pipeline.py of the reference (:12-78)."""
from afldm_b200.pipelines.i2sb_pipeline import I2SBLDMPipeline  # noqa: F401
