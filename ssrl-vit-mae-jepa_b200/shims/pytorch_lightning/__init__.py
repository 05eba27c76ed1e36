"""Minimal ``pytorch_lightning`` surface used by giolucasd/ssrl-vit-mae-jepa (only importable when the real package is not
installed: this directory sits at the end of ``sys.path``).

* ``LightningModule``: ``save_hyperparameters``, ``log``, ``current_epoch``, ``load_from_checkpoint`` (reference
  ``src/training/mae.py:14-23,49,56,83``, ``scripts/training/train_mae.py:85-91``)
* ``Trainer``: ``fit`` / ``test`` with the keyword arguments the reference passes (``scripts/training/pretrain_mae.py:116-128``,
  ``scripts/utils.py:186-199``), executed by ``ssrl_vit_mae_jepa_b200.trainer.LiteTrainer``
* ``seed_everything`` (``scripts/utils.py:19``)
"""
from __future__ import annotations

import inspect
import os
import random
from typing import Any, Dict, List, Optional

import torch

__version__ = "2.5.6"  # the release the reference pins (uv.lock:1772-1773); written into checkpoints
__shim__ = True


def seed_everything(seed: Optional[int] = None, workers: bool = False) -> int:
    seed = int(seed if seed is not None else os.environ.get("PL_GLOBAL_SEED", 0))
    os.environ["PL_GLOBAL_SEED"] = str(seed)
    os.environ["PL_SEED_WORKERS"] = str(int(workers))
    random.seed(seed)
    try:
        import numpy as np

        np.random.seed(seed % (2**32))
    except Exception:  # noqa: BLE001
        pass
    torch.manual_seed(seed)
    return seed


class LightningModule(torch.nn.Module):
    """The members of ``pl.LightningModule`` the reference's two training modules touch."""

    def __init__(self, *a, **k):
        super().__init__()
        self.current_epoch = 0
        self.logged: Dict[str, Any] = {}
        self.hparams: Dict[str, Any] = {}
        self.trainer = None

    def save_hyperparameters(self, *args, ignore=(), **kw):
        frame = inspect.currentframe().f_back
        skip = {"self", "__class__", *([ignore] if isinstance(ignore, str) else ignore)}
        self.hparams = {k: v for k, v in frame.f_locals.items() if k not in skip and not k.startswith("_")}

    def log(self, name: str, value, **kw):
        self.logged[name] = value

    def log_dict(self, d: Dict[str, Any], **kw):
        self.logged.update(d)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict: bool = True, **kwargs):
        ck = torch.load(str(checkpoint_path), map_location=map_location or "cpu", weights_only=False)
        hp = dict(ck.get("hyper_parameters", {}))
        hp.update(kwargs)
        accepted = set(inspect.signature(cls.__init__).parameters)
        module = cls(**{k: v for k, v in hp.items() if k in accepted})
        module.load_state_dict(ck["state_dict"], strict=strict)
        return module


class Trainer:
    """``pl.Trainer`` for one device (or one process per device under torchrun), backed by ``LiteTrainer``.
    ``precision="bf16-mixed"`` is what the kernels implement (bf16 tensor-core GEMMs, fp32 master weights / statistics);
    ``gradient_clip_val`` is clip-by-norm."""

    def __init__(self, accelerator: str = "auto", devices: Any = 1, max_epochs: Optional[int] = None, logger=None,
                 callbacks: Optional[List[Any]] = None, log_every_n_steps: int = 50, precision: Any = "32-true",
                 gradient_clip_val: Optional[float] = None, gradient_clip_algorithm: str = "norm",
                 limit_train_batches: Optional[int] = None, **kw):
        if gradient_clip_algorithm not in (None, "norm"):
            raise NotImplementedError("only clip-by-norm is implemented (the reference uses 'norm')")
        self.accelerator, self.devices, self.precision = accelerator, devices, precision
        self.max_epochs = 1000 if max_epochs is None else int(max_epochs)
        self.logger, self.callbacks = logger, list(callbacks or [])
        self.log_every_n_steps = log_every_n_steps
        self.gradient_clip_val = 0.0 if gradient_clip_val is None else float(gradient_clip_val)
        self.limit_batches = limit_train_batches if limit_train_batches is not None else (
            int(os.environ["MAE_LIMIT_BATCHES"]) if os.environ.get("MAE_LIMIT_BATCHES") else None)
        self._lite = None

    def _make(self):
        from ssrl_vit_mae_jepa_b200.trainer import LiteTrainer

        epochs = int(os.environ.get("MAE_MAX_EPOCHS", self.max_epochs))
        self._lite = LiteTrainer(max_epochs=epochs, callbacks=self.callbacks, gradient_clip_val=self.gradient_clip_val,
                                 log_every_n_steps=self.log_every_n_steps, limit_batches=self.limit_batches,
                                 logger=self.logger)
        return self._lite

    @property
    def current_epoch(self) -> int:
        return self._lite.current_epoch if self._lite else 0

    @property
    def global_step(self) -> int:
        return self._lite.global_step if self._lite else 0

    def fit(self, model, train_dataloaders=None, val_dataloaders=None, datamodule=None, ckpt_path: Optional[str] = None):
        model.trainer = self
        self._make().fit(model, train_dataloaders, val_dataloaders, ckpt_path=ckpt_path)

    def test(self, model, dataloaders=None, ckpt_path: Optional[str] = None, **kw) -> List[Dict[str, float]]:
        model.trainer = self
        return self._make().test(model, dataloaders, ckpt_path=ckpt_path)

    def validate(self, model, dataloaders=None, ckpt_path: Optional[str] = None, **kw) -> List[Dict[str, float]]:
        model.trainer = self
        return self._make().test(model, dataloaders, ckpt_path=ckpt_path, prefix="val")


from . import callbacks, loggers  # noqa: E402,F401
