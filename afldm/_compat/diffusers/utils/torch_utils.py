from afldm_b200.pipelines.ldm_pipeline import randn_tensor  # noqa: F401
