"""CPU oracle for the STL-10 input pipeline that precedes the MAE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/`` may import this module; the product (``ssrl-vit-mae-jepa_b200``) never does.

What this restates (integer / byte arithmetic -> numpy, bit-exact)
---------------------------------------------------------------
The reference's training transform ``/root/reference/src/data.py:15-25``::

    RandomResizedCrop(96, scale=(0.8, 1.0)) -> RandomHorizontalFlip() -> ToTensor() -> Normalize(.5, .5)

and its evaluation transform (``src/data.py:26-34``: ``Resize(96) -> CenterCrop(96)``, the identity on 96x96 STL-10
images, followed by the same ``ToTensor``/``Normalize``).  The arithmetic lives in third-party code that is not part
of ``/root/reference``: torchvision 0.26 (``transforms.RandomResizedCrop.get_params`` / ``functional.resized_crop``)
and Pillow 12.2 (``Image.resize(..., BILINEAR)`` = ``ImagingResample`` in ``src/libImaging/Resample.c``).  Their
published algorithms are restated here:

* ``get_params``: up to 10 attempts of ``area * U(s0, s1)`` and ``exp(U(log r0, log r1))``, ``w = round(sqrt(a*r))``,
  ``h = round(sqrt(a/r))``, accept if it fits, then ``i = randint(0, H-h+1)``, ``j = randint(0, W-w+1)``; centre-crop
  fallback.  RNG calls are made in the same order on the same torch generator, so a seeded run reproduces torchvision.
* Pillow bilinear resample of an 8-bit image: separable, horizontal pass then vertical pass, each pass with
  double-precision triangle-filter coefficients normalised per output pixel, converted to fixed point with
  ``PRECISION_BITS = 32 - 8 - 2 = 22`` (round half away from zero), accumulated in int32 starting from ``1 << 21`` and
  clipped to 8 bits BETWEEN the passes (the intermediate image is uint8).
* ``ToTensor`` = ``uint8 / 255`` in fp32, ``Normalize`` = ``(x - 0.5) / 0.5`` in fp32.

Pinning status: **pinned**.  Both third-party libraries are installed in this image, so ``tests/test_augment_oracle.py``
checks this restatement bit-for-bit against Pillow (``Image.crop().resize()``) and against the full torchvision
``Compose`` under a shared seed, and ``tests/golden/make_augment_golden.py`` writes fixtures produced by torchvision
itself for the GPU tests.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np
import torch

PRECISION_BITS = 32 - 8 - 2


def resample_coeffs(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Pillow ``precompute_coeffs`` + ``normalize_coeffs_8bpc`` for the bilinear (triangle) filter over the full box
    ``[0, in_size)``: returns ``(xmin (out,), count (out,), kk (out, ksize) int32)``."""
    scale = float(in_size) / float(out_size)
    filterscale = max(scale, 1.0)
    support = 1.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    xmin = np.zeros(out_size, np.int32)
    cnt = np.zeros(out_size, np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        x0 = int(center - support + 0.5)
        if x0 < 0:
            x0 = 0
        x1 = int(center + support + 0.5)
        if x1 > in_size:
            x1 = in_size
        n = x1 - x0
        w = np.zeros(n, np.float64)
        for x in range(n):
            t = (x + x0 - center + 0.5) * ss
            t = -t if t < 0.0 else t
            w[x] = 1.0 - t if t < 1.0 else 0.0
        ww = float(w.sum()) if n else 0.0
        # Pillow accumulates ww in the loop (left to right), identical to a sequential sum
        acc = 0.0
        for x in range(n):
            acc += w[x]
        ww = acc
        if ww != 0.0:
            w = w / ww
        for x in range(n):
            v = w[x] * (1 << PRECISION_BITS)
            kk[xx, x] = int(-0.5 + v) if w[x] < 0 else int(0.5 + v)
        xmin[xx], cnt[xx] = x0, n
    return xmin, cnt, kk


def _clip8(v: np.ndarray) -> np.ndarray:
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_bilinear_u8(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """Pillow ``Image.resize((out_w, out_h), BILINEAR)`` of an 8-bit image ``(H, W, C)`` (or ``(H, W)``)."""
    a = img if img.ndim == 3 else img[..., None]
    H, W, C = a.shape
    if out_w != W:
        xmin, cnt, kk = resample_coeffs(W, out_w)
        tmp = np.empty((H, out_w, C), np.uint8)
        for xx in range(out_w):
            acc = np.full((H, C), 1 << (PRECISION_BITS - 1), np.int64)
            for x in range(cnt[xx]):
                acc += a[:, xmin[xx] + x, :].astype(np.int64) * int(kk[xx, x])
            tmp[:, xx, :] = _clip8(acc)
        a = tmp
    if out_h != H:
        ymin, cnt, kk = resample_coeffs(H, out_h)
        out = np.empty((out_h, a.shape[1], C), np.uint8)
        for yy in range(out_h):
            acc = np.full((a.shape[1], C), 1 << (PRECISION_BITS - 1), np.int64)
            for y in range(cnt[yy]):
                acc += a[ymin[yy] + y, :, :].astype(np.int64) * int(kk[yy, y])
            out[yy] = _clip8(acc)
        a = out
    return a if img.ndim == 3 else a[..., 0]


def get_params(height: int, width: int, scale: Sequence[float], ratio: Sequence[float],
               generator: torch.Generator | None = None) -> Tuple[int, int, int, int]:
    """torchvision ``RandomResizedCrop.get_params`` with the same RNG call order (``generator=None`` = global RNG)."""
    area = height * width
    log_ratio = torch.log(torch.tensor(ratio))
    for _ in range(10):
        target_area = area * torch.empty(1).uniform_(scale[0], scale[1], generator=generator).item()
        aspect_ratio = torch.exp(torch.empty(1).uniform_(log_ratio[0], log_ratio[1], generator=generator)).item()
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = torch.randint(0, height - h + 1, size=(1,), generator=generator).item()
            j = torch.randint(0, width - w + 1, size=(1,), generator=generator).item()
            return i, j, h, w
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w


def get_params_from_uniforms(height: int, width: int, scale: Sequence[float], ratio: Sequence[float],
                             u: np.ndarray) -> Tuple[int, int, int, int, int]:
    """``get_params`` (above) + the flip coin with the random numbers made explicit: ``u`` holds 23 fp32 uniforms in
    [0, 1): ``u[2a], u[2a+1]`` = area / log-ratio draw of attempt ``a``, ``u[20], u[21]`` = top / left, ``u[22]`` = flip.
    Same arithmetic as torchvision: fp32 ``uniform_`` (``u * (to - from) + from``) and fp32 ``exp``, Python doubles
    and ``round`` after ``.item()``.  Oracle of ``mae_augment_params`` (whose generator is counter-based, so the uniforms
    are reproducible on the host)."""
    area = height * width
    s0, s1 = np.float32(scale[0]), np.float32(scale[1])
    l0, l1 = np.log(np.float32(ratio[0])), np.log(np.float32(ratio[1]))
    flip = int(np.float32(u[22]) < np.float32(0.5))
    for a in range(10):
        sf = np.float32(np.float32(u[2 * a]) * np.float32(s1 - s0)) + s0
        lr = np.float32(np.float32(u[2 * a + 1]) * np.float32(l1 - l0)) + l0
        target_area = area * float(np.float32(sf))
        aspect_ratio = float(np.exp(np.float32(lr)))
        w = int(round(math.sqrt(target_area * aspect_ratio)))
        h = int(round(math.sqrt(target_area / aspect_ratio)))
        if 0 < w <= width and 0 < h <= height:
            i = int(float(np.float32(u[20])) * (height - h + 1))
            j = int(float(np.float32(u[21])) * (width - w + 1))
            return i, j, h, w, flip
    in_ratio = float(width) / float(height)
    if in_ratio < min(ratio):
        w = width
        h = int(round(w / min(ratio)))
    elif in_ratio > max(ratio):
        h = height
        w = int(round(h * max(ratio)))
    else:
        w, h = width, height
    return (height - h) // 2, (width - w) // 2, h, w, flip


def splitmix64(x: np.ndarray) -> np.ndarray:
    """The 64-bit mixer of the device generator (uint64 arrays, wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
        return x ^ (x >> np.uint64(31))


def device_uniforms(n: int, seed: int, counter: int) -> np.ndarray:
    """(n, 23) fp32 uniforms exactly as ``mae_augment_params`` generates them for samples 0..n-1."""
    with np.errstate(over="ignore"):
        key = splitmix64(splitmix64(np.array([seed + counter], np.uint64)) + np.arange(n, dtype=np.uint64))
        v = splitmix64(key[:, None] + np.arange(23, dtype=np.uint64)[None, :])
    return ((v >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 16777216.0)).astype(np.float32)


def train_params(n: int, size: int = 96, scale=(0.8, 1.0), ratio=(3.0 / 4.0, 4.0 / 3.0),
                 generator: torch.Generator | None = None) -> np.ndarray:
    """Per-sample ``(i, j, h, w, flip)`` int32 rows in the order the reference's Compose consumes the RNG:
    ``get_params`` then ``torch.rand(1) < 0.5`` for the flip."""
    out = np.zeros((n, 5), np.int32)
    for k in range(n):
        i, j, h, w = get_params(size, size, scale, ratio, generator)
        flip = int(torch.rand(1, generator=generator).item() < 0.5)
        out[k] = (i, j, h, w, flip)
    return out


def to_tensor_normalize(u8_hwc: np.ndarray) -> np.ndarray:
    """``ToTensor`` + ``Normalize((.5,.5,.5), (.5,.5,.5))``: ``(H, W, C) uint8 -> (C, H, W) float32``."""
    x = u8_hwc.astype(np.float32) / np.float32(255.0)
    x = (x - np.float32(0.5)) / np.float32(0.5)
    return np.ascontiguousarray(x.transpose(2, 0, 1))


def augment_batch(images_u8_nchw: np.ndarray, params: np.ndarray, out_size: int = 96) -> np.ndarray:
    """The whole transform for a batch stored like ``STL10.data``: ``(N, 3, H, W) uint8`` -> ``(N, 3, S, S) float32``.
    ``params[k] = (i, j, h, w, flip)``; the evaluation transform is ``(0, 0, H, W, 0)``."""
    n = images_u8_nchw.shape[0]
    out = np.empty((n, images_u8_nchw.shape[1], out_size, out_size), np.float32)
    for k in range(n):
        i, j, h, w, flip = [int(v) for v in params[k]]
        hwc = np.ascontiguousarray(images_u8_nchw[k].transpose(1, 2, 0))
        crop = hwc[i:i + h, j:j + w]
        res = resize_bilinear_u8(crop, out_size, out_size)
        if flip:
            res = res[:, ::-1]
        out[k] = to_tensor_normalize(res)
    return out
