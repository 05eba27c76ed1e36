"""STL-10 input pipeline with the reference's entry points (``src/data.py:45-185``: ``get_pretrain_dataloaders``,
``get_train_dataloaders``, ``get_test_dataloader``) and transforms (RandomResizedCrop(96, 0.8-1.0) + flip for
training, Resize/CenterCrop for evaluation, ToTensor, Normalize(0.5, 0.5)).

STL-10 cannot be downloaded in this environment (no network), so when the dataset is not already on disk under
``data/`` -- or when ``MAE_SYNTHETIC_DATA=1`` -- a deterministic synthetic STL-10-shaped dataset (same sizes,
value range [-1, 1] after normalisation, 10 classes) is used instead; everything downstream is unchanged.
"""
from __future__ import annotations

import os
from typing import Tuple

import torch
from torch.utils.data import DataLoader, Dataset, Subset, random_split

DATA_DIR = os.environ.get("MAE_DATA_DIR", "data")
SPLIT_SIZES = {"unlabeled": 100_000, "train": 5_000, "test": 8_000}


class SyntheticSTL10(Dataset):
    """Deterministic stand-in: item i is a smooth random 96x96 RGB field in [-1, 1] (seeded by (split, i)) with label
    i % 10.  ``train=True`` applies a random horizontal flip like the reference's augmentation."""

    def __init__(self, split: str, train: bool = True, size: int | None = None, image_size: int = 96):
        self.split, self.train, self.image_size = split, train, image_size
        self.n = int(size if size is not None else SPLIT_SIZES[split])
        self.salt = {"unlabeled": 0, "train": 1, "test": 2}[split]

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        g = torch.Generator().manual_seed(self.salt * 1_000_003 + int(i))
        low = torch.rand(3, 12, 12, generator=g)
        img = torch.nn.functional.interpolate(low[None], size=self.image_size, mode="bilinear", align_corners=False)[0]
        img = (img + 0.1 * torch.rand(3, self.image_size, self.image_size, generator=g)).clamp_(0, 1)
        if self.train and torch.rand(1, generator=g).item() < 0.5:
            img = img.flip(-1)
        label = int(i) % 10 if self.split != "unlabeled" else -1
        return (img - 0.5) / 0.5, label


def _real_stl10(split: str, train: bool):
    if os.environ.get("MAE_SYNTHETIC_DATA") == "1":
        return None
    try:
        from torchvision import transforms
        from torchvision.datasets import STL10

        if train:
            tf = transforms.Compose([transforms.RandomResizedCrop(96, scale=(0.8, 1.0)), transforms.RandomHorizontalFlip(),
                                     transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
        else:
            tf = transforms.Compose([transforms.Resize(96), transforms.CenterCrop(96), transforms.ToTensor(),
                                     transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
        return STL10(DATA_DIR, split=split, transform=tf, download=False)
    except Exception:  # noqa: BLE001 - dataset not on disk / torchvision missing -> synthetic
        return None


def _dataset(split: str, train: bool, synthetic_size: int | None = None):
    ds = _real_stl10(split, train)
    if ds is None:
        size = synthetic_size or int(os.environ.get("MAE_SYNTHETIC_SIZE", SPLIT_SIZES[split]))
        ds = SyntheticSTL10(split, train=train, size=size)
        print(f"[data] STL-10 '{split}' split not found under '{DATA_DIR}': using {len(ds)} synthetic STL-10-shaped samples")
    return ds


def _loader(ds, batch_size: int, shuffle: bool, num_workers: int) -> DataLoader:
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, pin_memory=True,
                      drop_last=False, persistent_workers=num_workers > 0)


def get_pretrain_dataloaders(cfg: dict) -> Tuple[DataLoader, DataLoader]:
    """Unlabeled split -> (train, val) loaders; keys ``pretrain.{batch_size,val_split,data_fraction,num_workers}``, ``seed``."""
    pre, seed = cfg["pretrain"], cfg.get("seed", 73)
    full = _dataset("unlabeled", train=True)
    frac = pre.get("data_fraction", 1.0)
    if frac < 1.0:
        full = Subset(full, range(int(len(full) * frac)))
    n_val = int(len(full) * pre.get("val_split", 0.1))
    train_ds, val_ds = random_split(full, [len(full) - n_val, n_val], generator=torch.Generator().manual_seed(seed))
    bs, nw = pre.get("batch_size", 512), pre.get("num_workers", 4)
    print(f"[data] unlabeled pretrain split: {len(train_ds)} train, {n_val} val")
    return _loader(train_ds, bs, True, nw), _loader(val_ds, bs, False, nw)


def get_train_dataloaders(cfg: dict) -> Tuple[DataLoader, DataLoader]:
    """Labeled train split -> (train, val): ``train.samples_per_class`` images per class for training, the rest for
    validation (reference ``src/data.py:109-164``)."""
    tr, seed = cfg["train"], cfg.get("seed", 73)
    full = _dataset("train", train=True)
    per_class = int(tr.get("samples_per_class", 400))
    g = torch.Generator().manual_seed(seed)
    order = torch.randperm(len(full), generator=g).tolist()
    labels = [full[i][1] if isinstance(full, SyntheticSTL10) else int(full.labels[i]) for i in order]
    taken, train_idx, val_idx = {}, [], []
    for i, y in zip(order, labels):
        if taken.get(y, 0) < per_class:
            taken[y] = taken.get(y, 0) + 1
            train_idx.append(i)
        else:
            val_idx.append(i)
    bs, nw = tr.get("batch_size", 512), tr.get("num_workers", 4)
    print(f"[data] labeled split: {len(train_idx)} train, {len(val_idx)} val")
    return _loader(Subset(full, train_idx), bs, True, nw), _loader(Subset(full, val_idx or train_idx[:1]), bs, False, nw)


def get_test_dataloader(cfg: dict) -> DataLoader:
    te = cfg.get("test", {})
    return _loader(_dataset("test", train=False), te.get("batch_size", 512), False, te.get("num_workers", 4))
