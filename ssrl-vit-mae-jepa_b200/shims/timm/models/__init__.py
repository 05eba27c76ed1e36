from . import vision_transformer  # noqa: F401
