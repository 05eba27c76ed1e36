"""Same import path as the reference (``from src.training.mae import MAEPretrainModule``,
reference ``scripts/training/pretrain_mae.py:13``): re-exports the B200-native implementation."""
from ssrl_vit_mae_jepa_b200.training import MAEPretrainModule  # noqa: F401
