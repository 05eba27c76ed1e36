"""``pytorch_lightning.loggers.TensorBoardLogger`` (reference ``scripts/training/pretrain_mae.py:83``): scalars go to
``<save_dir>/<name>/version_<n>`` through ``torch.utils.tensorboard`` when tensorboard is installed, else to a CSV file."""
from __future__ import annotations

import os
from typing import Dict, Optional


class TensorBoardLogger:
    def __init__(self, save_dir: str, name: Optional[str] = "lightning_logs", version: Optional[int] = None, **kw):
        self.save_dir, self.name = str(save_dir), name or ""
        root = os.path.join(self.save_dir, self.name)
        if version is None:
            existing = [d for d in (os.listdir(root) if os.path.isdir(root) else []) if d.startswith("version_")]
            version = 1 + max([int(d.split("_")[1]) for d in existing if d.split("_")[1].isdigit()], default=-1)
        self.version = version
        self._writer = None
        self._csv = None

    @property
    def log_dir(self) -> str:
        return os.path.join(self.save_dir, self.name, f"version_{self.version}")

    def log_metrics(self, metrics: Dict[str, float], step: int) -> None:
        os.makedirs(self.log_dir, exist_ok=True)
        if self._writer is None and self._csv is None:
            try:
                from torch.utils.tensorboard import SummaryWriter

                self._writer = SummaryWriter(self.log_dir)
            except Exception:  # noqa: BLE001 - tensorboard not installed
                self._csv = open(os.path.join(self.log_dir, "metrics.csv"), "a")
        for k, v in metrics.items():
            if self._writer is not None:
                self._writer.add_scalar(k, float(v), step)
            else:
                self._csv.write(f"{step},{k},{float(v)}\n")
        if self._csv is not None:
            self._csv.flush()

    def finalize(self) -> None:
        if self._writer is not None:
            self._writer.flush()
