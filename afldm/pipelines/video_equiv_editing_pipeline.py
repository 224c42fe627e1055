"""afldm/pipelines/video_equiv_editing_pipeline.py of the reference (:330-748)."""
from afldm_b200.pipelines.video_equiv_editing_pipeline import VideoEquivariantEditingPipeline  # noqa: F401
