"""Minimal ``timm`` surface used by giolucasd/ssrl-vit-mae-jepa: ``timm.models.vision_transformer.VisionTransformer``
(reference ``scripts/training/train_mae.py:11,152-160``, ``scripts/utils.py:7``), backed by the B200 kernels."""
__version__ = "1.0.21"  # the release the reference pins (uv.lock:2163-2164)
__shim__ = True
from . import models  # noqa: E402,F401
