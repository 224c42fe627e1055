from . import torch_utils  # noqa: F401


def load_image(image, convert_method=None):
    """diffusers.utils.load_image: path / PIL image -> RGB PIL image (EXIF orientation applied)."""
    import PIL.Image
    import PIL.ImageOps
    if isinstance(image, str):
        if image.startswith(("http://", "https://")):
            raise ValueError("no network access: pass a local file")
        image = PIL.Image.open(image)
    elif not isinstance(image, PIL.Image.Image):
        raise ValueError("Incorrect format used for the image. Should be a local path or a PIL image.")
    image = PIL.ImageOps.exif_transpose(image)
    return convert_method(image) if convert_method is not None else image.convert("RGB")


def export_to_video(*args, **kwargs):
    raise NotImplementedError("export_to_video: video file output is outside the AF-LDM hot path of this build")
