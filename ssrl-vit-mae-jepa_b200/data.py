"""STL-10 input pipeline with the reference's entry points (``src/data.py:45-185``: ``get_pretrain_dataloaders``,
``get_train_dataloaders``, ``get_test_dataloader``) and transforms (RandomResizedCrop(96, 0.8-1.0) + flip for
training, Resize/CenterCrop for evaluation, ToTensor, Normalize(0.5, 0.5)).

STL-10 is read from ``MAE_DATA_DIR`` (default ``data/``) and downloaded on first use like the reference.  This
environment has no network, so the tests and benchmarks opt in to a deterministic synthetic STL-10-shaped dataset (same
sizes, value range [-1, 1] after normalisation, 10 classes) with ``MAE_SYNTHETIC_DATA=1``; without that variable a missing
dataset raises (no silent training on noise).

``<section>.gpu_augment: true`` in the YAML switches the loaders to raw uint8 batches; the trainer then applies the same
transforms on the GPU (``GpuAugment`` -> ``mae_augment_u8``, bit-exact Pillow bilinear), which removes the CPU
transform from the loader workers and cuts the host->device copy 4x.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import torch
from torch.utils.data import DataLoader, Dataset, Subset

DATA_DIR = os.environ.get("MAE_DATA_DIR", "data")
SPLIT_SIZES = {"unlabeled": 100_000, "train": 5_000, "test": 8_000}


class SyntheticSTL10(Dataset):
    """Deterministic stand-in: item i is a smooth random 96x96 RGB field in [-1, 1] (seeded by (split, i)) with label
    i % 10.  ``train=True`` applies a random horizontal flip like the reference's augmentation."""

    def __init__(self, split: str, train: bool = True, size: int | None = None, image_size: int = 96,
                 raw_uint8: bool = False):
        self.split, self.train, self.image_size, self.raw_uint8 = split, train, image_size, raw_uint8
        self.n = int(size if size is not None else SPLIT_SIZES[split])
        self.salt = {"unlabeled": 0, "train": 1, "test": 2}[split]

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int):
        g = torch.Generator().manual_seed(self.salt * 1_000_003 + int(i))
        low = torch.rand(3, 12, 12, generator=g)
        img = torch.nn.functional.interpolate(low[None], size=self.image_size, mode="bilinear", align_corners=False)[0]
        img = (img + 0.1 * torch.rand(3, self.image_size, self.image_size, generator=g)).clamp_(0, 1)
        if self.raw_uint8:  # as stored by torchvision's STL10.data; crop / flip / normalize happen on the GPU (GpuAugment)
            return (img * 255.0).round().to(torch.uint8), (int(i) % 10 if self.split != "unlabeled" else -1)
        if self.train and torch.rand(1, generator=g).item() < 0.5:
            img = img.flip(-1)
        label = int(i) % 10 if self.split != "unlabeled" else -1
        return (img - 0.5) / 0.5, label


class RawSTL10(Dataset):
    """The real STL-10 arrays (``STL10.data`` (N,3,96,96) uint8, ``STL10.labels``) without any CPU transform."""

    def __init__(self, stl10):
        self.data, self.labels = stl10.data, getattr(stl10, "labels", None)

    def __len__(self) -> int:
        return len(self.data)

    def __getitem__(self, i: int):
        return torch.from_numpy(self.data[i]), (int(self.labels[i]) if self.labels is not None else -1)


class GpuAugment:
    """The reference's transforms (``src/data.py:15-34``) applied on the GPU to uint8 batches (B,3,H,W):
    ``train=True``: RandomResizedCrop(size, scale, ratio) -> RandomHorizontalFlip -> ToTensor -> Normalize(.5,.5);
    ``train=False``: Resize/CenterCrop (identity for 96x96 STL-10) -> ToTensor -> Normalize.
    The pixels come from ``mae_augment_u8`` (bit-exact Pillow bilinear); the random crop boxes / flips follow
    ``RandomResizedCrop.get_params`` (10 attempts of area ~ U(scale), log-ratio ~ U(log ratio), first fitting one wins,
    centre-crop fallback) and are drawn on the device by ``mae_augment_params`` (counter-based generator: nothing of the
    input pipeline runs on the host per step); passing a ``torch.Generator`` switches to drawing them with torch instead.
    H2D traffic is 4x smaller than shipping float32 images."""

    def __init__(self, size: int = 96, scale=(0.8, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0), train: bool = True,
                 generator: Optional[torch.Generator] = None, seed: int = 73):
        self.size, self.scale, self.ratio, self.train, self.generator = size, scale, ratio, train, generator
        self._next, self._next_key = None, None  # (torch-generator mode) parameters drawn ahead for the next call
        # default: boxes / flips drawn on the device by mae_augment_params, a pure function of (seed, call counter, sample)
        self.seed, self.counter = int(seed), 0
        self._params: Optional[torch.Tensor] = None

    def device_params(self, B: int, H: int, W: int, device) -> torch.Tensor:
        """(B,5) int32 rows (i, j, h, w, flip) drawn on ``device`` (no host RNG, no host <-> device copy, no sync)."""
        from ._lib import calls

        if self._params is None or self._params.shape[0] != B or self._params.device != device:
            self._params = torch.empty(B, 5, dtype=torch.int32, device=device)
        calls.mae_augment_params(self._params.data_ptr(), B, H, W, float(self.scale[0]), float(self.scale[1]),
                                 float(self.ratio[0]), float(self.ratio[1]), self.seed, self.counter, int(self.train),
                                 torch.cuda.current_stream(device).cuda_stream)
        self.counter += 1
        return self._params

    def draw_params(self, B: int, H: int, W: int, device) -> torch.Tensor:
        """(B,5) int32 rows (i, j, h, w, flip) on ``device``."""
        if not self.train:
            p = torch.zeros(B, 5, dtype=torch.int32, device=device)
            p[:, 2], p[:, 3] = H, W
            return p
        g = self.generator
        u = torch.rand(B, 10, 2, device=device, generator=g, dtype=torch.float32)
        area = (H * W) * (self.scale[0] + (self.scale[1] - self.scale[0]) * u[..., 0]).double()
        lo, hi = math.log(self.ratio[0]), math.log(self.ratio[1])
        aspect = torch.exp((lo + (hi - lo) * u[..., 1])).double()
        w = torch.round(torch.sqrt(area * aspect)).long()
        h = torch.round(torch.sqrt(area / aspect)).long()
        ok = (w > 0) & (w <= W) & (h > 0) & (h <= H)
        first = torch.argmax(ok.int(), dim=1, keepdim=True)  # first fitting attempt (0 if none)
        any_ok = ok.any(dim=1)
        w = torch.where(any_ok, w.gather(1, first)[:, 0], torch.full_like(any_ok, W, dtype=torch.long))
        h = torch.where(any_ok, h.gather(1, first)[:, 0], torch.full_like(any_ok, H, dtype=torch.long))
        # fallback for an image whose aspect is inside `ratio` (square STL-10): the whole image
        v = torch.rand(B, 3, device=device, generator=g, dtype=torch.float32).double()
        i = torch.clamp((v[:, 0] * (H - h + 1)).long(), max=H - 1)
        j = torch.clamp((v[:, 1] * (W - w + 1)).long(), max=W - 1)
        i, j = torch.minimum(i, H - h), torch.minimum(j, W - w)
        flip = (v[:, 2] < 0.5).long()
        return torch.stack([i, j, h, w, flip], dim=1).to(torch.int32).contiguous()

    def __call__(self, images_u8: torch.Tensor, params: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None) -> torch.Tensor:
        from ._lib import calls

        if not images_u8.is_cuda or images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
            raise RuntimeError("GpuAugment needs a CUDA uint8 (B,C,H,W) batch (there is no CPU fallback)")
        images_u8 = images_u8.contiguous()
        B, C, H, W = images_u8.shape
        drawn_here = params is None and self.generator is not None
        if params is None and self.generator is None:
            params = self.device_params(B, H, W, images_u8.device)
        elif drawn_here:
            # torch-generator mode (reproducing a torch RNG stream): drawn on the host with a few vectorised ops and shipped
            # as one small copy from a pinned buffer; the draw for the NEXT call is made right after this call's kernel launch
            key = (B, H, W)
            if self._next_key != key:
                self._next, self._next_key = self._draw_host(B, H, W), key
            params = self._next
        params = params.to(device=images_u8.device, dtype=torch.int32, non_blocking=True).contiguous()
        if params.shape != (B, 5):
            raise ValueError(f"params must be ({B},5)")
        if out is None:
            out = torch.empty(B, C, self.size, self.size, dtype=torch.float32, device=images_u8.device)
        elif out.shape != (B, C, self.size, self.size) or out.dtype != torch.float32 or not out.is_contiguous() or \
                out.device != images_u8.device:
            raise ValueError(f"out must be a contiguous float32 ({B},{C},{self.size},{self.size}) tensor on the input's device")
        calls.mae_augment_u8(images_u8.data_ptr(), params.data_ptr(), B, C, H, W, self.size, out.data_ptr(),
                             torch.cuda.current_stream(images_u8.device).cuda_stream)
        if drawn_here:
            self._next = self._draw_host(B, H, W)  # the previous pinned buffer is kept alive by `params` until its copy ran
        return out

    def _draw_host(self, B: int, H: int, W: int) -> torch.Tensor:
        where = self.generator.device if self.generator is not None else torch.device("cpu")
        p = self.draw_params(B, H, W, where)
        return p.pin_memory() if p.device.type == "cpu" else p


def _real_stl10(split: str, train: bool, raw_uint8: bool = False):
    """torchvision STL-10 with the reference's transforms, downloaded on first use exactly like the reference
    (``src/data.py:60-65``).  The synthetic stand-in is used ONLY on the explicit opt-in ``MAE_SYNTHETIC_DATA=1``: a wrong
    ``MAE_DATA_DIR``, a corrupt archive or a missing torchvision raises instead of silently training on noise."""
    if os.environ.get("MAE_SYNTHETIC_DATA") == "1":
        return None
    try:
        from torchvision import transforms
        from torchvision.datasets import STL10

        if raw_uint8:
            return RawSTL10(STL10(DATA_DIR, split=split, transform=None, download=True))
        if train:
            tf = transforms.Compose([transforms.RandomResizedCrop(96, scale=(0.8, 1.0)), transforms.RandomHorizontalFlip(),
                                     transforms.ToTensor(), transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
        else:
            tf = transforms.Compose([transforms.Resize(96), transforms.CenterCrop(96), transforms.ToTensor(),
                                     transforms.Normalize((0.5,) * 3, (0.5,) * 3)])
        return STL10(DATA_DIR, split=split, transform=tf, download=True)
    except Exception as e:  # noqa: BLE001
        raise RuntimeError(f"STL-10 split '{split}' is not available under '{DATA_DIR}' and could not be downloaded ({e!r}); "
                           "point MAE_DATA_DIR at the dataset, or set MAE_SYNTHETIC_DATA=1 to train on the synthetic "
                           "STL-10-shaped stand-in on purpose") from e


def _dataset(split: str, train: bool, synthetic_size: int | None = None, raw_uint8: bool = False):
    ds = _real_stl10(split, train, raw_uint8)
    if ds is None:
        size = synthetic_size or int(os.environ.get("MAE_SYNTHETIC_SIZE", SPLIT_SIZES[split]))
        ds = SyntheticSTL10(split, train=train, size=size, raw_uint8=raw_uint8)
        print(f"[data] MAE_SYNTHETIC_DATA=1: '{split}' split = {len(ds)} synthetic STL-10-shaped samples")
    return ds


def _labels_of(ds) -> "np.ndarray":
    import numpy as np

    if isinstance(ds, SyntheticSTL10):
        return np.arange(len(ds)) % 10
    return np.asarray(ds.labels)


def _loader(ds, batch_size: int, shuffle: bool, num_workers: int) -> DataLoader:
    return DataLoader(ds, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, pin_memory=True,
                      drop_last=False, persistent_workers=num_workers > 0)


def get_pretrain_dataloaders(cfg: dict) -> Tuple[DataLoader, DataLoader]:
    """Unlabeled split -> (train, val) loaders; keys ``pretrain.{batch_size,val_split,data_fraction,num_workers}``, ``seed``
    (reference ``src/data.py:45-106``).  The validation subset reads the SAME indices through a second dataset object
    carrying the evaluation transform (the reference ends up validating with the eval transform, ``src/data.py:81``),
    while training keeps RandomResizedCrop + flip."""
    pre, seed = cfg["pretrain"], cfg.get("seed", 73)
    raw = bool(pre.get("gpu_augment", False))
    full_train = _dataset("unlabeled", train=True, raw_uint8=raw)
    full_eval = full_train if raw else _dataset("unlabeled", train=False, raw_uint8=raw)  # raw uint8: the trainer picks the transform
    n_full = len(full_train)
    frac = pre.get("data_fraction", 1.0)
    n = int(n_full * frac) if frac < 1.0 else n_full
    n_val = int(n * pre.get("val_split", 0.1))
    perm = torch.randperm(n, generator=torch.Generator().manual_seed(seed)).tolist()  # == random_split's permutation
    train_ds, val_ds = Subset(full_train, perm[:n - n_val]), Subset(full_eval, perm[n - n_val:])
    bs, nw = pre.get("batch_size", 512), pre.get("num_workers", 4)
    print(f"[data] unlabeled pretrain split: {len(train_ds)} train, {n_val} val")
    return _loader(train_ds, bs, True, nw), _loader(val_ds, bs, False, nw)


def get_train_dataloaders(cfg: dict) -> Tuple[DataLoader, DataLoader]:
    """Labeled train split -> (train, val): per class, the indices are shuffled with ``numpy.random.default_rng(seed)``
    and the first ``train.samples_per_class`` go to training, the rest to validation -- the same membership as the
    reference (``src/data.py:128-139``); validation reads through the evaluation transform."""
    import numpy as np

    tr, seed = cfg["train"], cfg.get("seed", 73)
    raw = bool(tr.get("gpu_augment", False))
    full_train = _dataset("train", train=True, raw_uint8=raw)
    full_eval = full_train if raw else _dataset("train", train=False, raw_uint8=raw)
    labels = _labels_of(full_train)
    per_class = int(tr.get("samples_per_class", 400))
    train_idx, val_idx = [], []
    for c in np.unique(labels):
        cls_idx = np.where(labels == c)[0]
        np.random.default_rng(seed).shuffle(cls_idx)
        train_idx.extend(int(i) for i in cls_idx[:per_class])
        val_idx.extend(int(i) for i in cls_idx[per_class:])
    bs, nw = tr.get("batch_size", 256), tr.get("num_workers", 4)
    print(f"[data] labeled split: {len(train_idx)} train, {len(val_idx)} val")
    return _loader(Subset(full_train, train_idx), bs, True, nw), _loader(Subset(full_eval, val_idx or train_idx[:1]), bs, False, nw)


def get_test_dataloader(cfg: dict) -> DataLoader:
    te = cfg.get("test", {})
    return _loader(_dataset("test", train=False, raw_uint8=bool(te.get("gpu_augment", False))), te.get("batch_size", 256), False,
                   te.get("num_workers", 4))
