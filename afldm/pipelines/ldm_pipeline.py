"""afldm/pipelines/ldm_pipeline.py of the reference (:32-160)."""
from afldm_b200.pipelines.ldm_pipeline import ImagePipelineOutput, MyLDMPipeline  # noqa: F401
