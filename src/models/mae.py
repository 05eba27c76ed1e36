"""Same import path as the reference (``from src.models.mae import MaskedAutoencoder``,
reference ``src/training/mae.py:11``): re-exports the B200-native implementation."""
from ssrl_vit_mae_jepa_b200.mae import MaskedAutoencoder  # noqa: F401
