"""``pytorch_lightning.callbacks`` names the reference imports (``scripts/training/pretrain_mae.py:9``)."""
from __future__ import annotations


class LearningRateMonitor:
    """The learning rate is logged by the trainer itself (``lr`` scalar per epoch); this object only carries the option."""

    def __init__(self, logging_interval: str = "epoch", **kw):
        self.logging_interval = logging_interval


def __getattr__(name):
    if name == "ModelCheckpoint":  # lazy: the trainer module imports torch + the CUDA package
        from ssrl_vit_mae_jepa_b200.trainer import ModelCheckpoint

        return ModelCheckpoint
    raise AttributeError(name)
