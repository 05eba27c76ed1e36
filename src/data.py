"""Same import path as the reference (``from src.data import get_pretrain_dataloaders``)."""
from ssrl_vit_mae_jepa_b200.data import get_pretrain_dataloaders, get_test_dataloader, get_train_dataloaders  # noqa: F401
