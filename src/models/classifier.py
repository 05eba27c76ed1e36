"""Same import path as the reference (``from src.models.classifier import ViTClassifier``): B200-native implementation."""
from ssrl_vit_mae_jepa_b200.classifier import ClassificationHead, ViTClassifier  # noqa: F401
