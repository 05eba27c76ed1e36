"""Same import path as the reference (``from src.training.classifier import ViTClassifierTrainModule``)."""
from ssrl_vit_mae_jepa_b200.classifier import ViTClassifierTrainModule  # noqa: F401
