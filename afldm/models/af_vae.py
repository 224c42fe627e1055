"""afldm/models/af_vae.py of the reference (:8-55)."""
from afldm_b200.models.af_vae import AliasFreeAutoencoderKL  # noqa: F401
