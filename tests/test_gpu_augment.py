"""-m gpu: ``mae_augment_u8`` (through the C ABI) against the pinned oracle and the torchvision-made golden vectors.
Integer path: bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A
from ssrl_vit_mae_jepa_b200._lib import calls
from ssrl_vit_mae_jepa_b200.data import GpuAugment
from tests.gpu_util import p, stream

pytestmark = pytest.mark.gpu
dev = "cuda"
GOLD = os.path.join(os.path.dirname(__file__), "golden", "augment_seed73.npz")


def _run(images_u8, params, S=96):
    x = torch.from_numpy(np.ascontiguousarray(images_u8)).to(dev)
    pr = torch.from_numpy(np.ascontiguousarray(params.astype(np.int32))).to(dev)
    B, C, H, W = x.shape
    out = torch.empty(B, C, S, S, device=dev)
    calls.mae_augment_u8(p(x), p(pr), B, C, H, W, S, p(out), stream())
    return out.cpu().numpy()


def test_golden_torchvision_vectors_bit_exact():
    g = np.load(GOLD)
    got = _run(g["images_u8"], g["params"])
    assert np.array_equal(got, g["expected"])


def test_random_and_edge_boxes_vs_oracle():
    rng = np.random.default_rng(3)
    B = 48
    imgs = rng.integers(0, 256, (B, 3, 96, 96), dtype=np.uint8)
    imgs[0] = 255; imgs[1] = 0  # saturated images: the clip8 edges
    torch.manual_seed(11)
    params = A.train_params(B)
    edge = [(0, 0, 96, 96, 0), (0, 0, 96, 96, 1), (0, 0, 1, 1, 0), (95, 95, 1, 1, 1), (10, 0, 74, 96, 1), (0, 22, 96, 74, 0),
            (48, 48, 33, 17, 0), (3, 5, 93, 91, 1)]
    params[: len(edge)] = np.array(edge, np.int32)
    got = _run(imgs, params)
    ref = A.augment_batch(imgs, params)
    assert np.array_equal(got, ref)


def test_non_square_source_and_smaller_output():
    rng = np.random.default_rng(4)
    imgs = rng.integers(0, 256, (5, 1, 80, 120), dtype=np.uint8)
    params = np.array([(0, 0, 80, 120, 0), (5, 10, 60, 64, 1), (0, 0, 64, 64, 0), (79, 119, 1, 1, 0), (8, 8, 40, 100, 1)], np.int32)
    got = _run(imgs, params, S=64)  # includes 120 -> 64 and 100 -> 64: anti-aliased (wider support) down-scaling
    ref = A.augment_batch(imgs, params, out_size=64)
    assert np.array_equal(got, ref)


def test_full_batch_properties():
    """B = 2000 (BASELINE.json's batch): size-independent properties instead of the slow oracle."""
    B = 2000
    g = torch.Generator(device=dev).manual_seed(5)
    x = torch.randint(0, 256, (B, 3, 96, 96), dtype=torch.uint8, device=dev, generator=g)
    ev = GpuAugment(train=False)(x)
    # evaluation transform = ToTensor + Normalize exactly (CPU arithmetic: torch's CUDA `/ 255` multiplies by a reciprocal)
    assert torch.equal(ev.cpu(), (x.cpu().float() / 255.0 - 0.5) / 0.5)
    aug = GpuAugment(train=True, generator=g)
    pr = aug.draw_params(B, 96, 96, dev)
    assert pr.shape == (B, 5) and pr.dtype == torch.int32
    i, j, h, w, f = [pr[:, k].long() for k in range(5)]
    assert (h >= 74).all() and (h <= 96).all() and (w >= 74).all() and (w <= 96).all()
    assert (i >= 0).all() and (j >= 0).all() and (i + h <= 96).all() and (j + w <= 96).all()
    assert 0.4 < f.float().mean().item() < 0.6 and 0.78 * 9216 < (h * w).float().mean().item() < 0.95 * 9216
    out = aug(x, pr)
    assert out.min() >= -1.0 and out.max() <= 1.0
    pr2 = pr.clone(); pr2[:, 4] = 1 - pr2[:, 4]
    assert torch.equal(aug(x, pr2), out.flip(-1))  # flipping commutes with everything after the resample
    # a handful of samples against the oracle
    sel = [0, 7, 999, 1999]
    ref = A.augment_batch(x[sel].cpu().numpy(), pr[sel].cpu().numpy())
    assert np.array_equal(out[sel].cpu().numpy(), ref)


def test_rejects_cpu_and_bad_shapes():
    with pytest.raises(RuntimeError):
        GpuAugment()(torch.zeros(2, 3, 96, 96, dtype=torch.uint8))
    x = torch.zeros(2, 3, 96, 96, dtype=torch.uint8, device=dev)
    with pytest.raises(ValueError):
        GpuAugment()(x, torch.zeros(3, 5, dtype=torch.int32, device=dev))
    with pytest.raises(RuntimeError):  # source more than 3x the output: beyond the kernel's tap budget
        out = torch.empty(1, 1, 8, 8, device=dev)
        big = torch.zeros(1, 1, 96, 96, dtype=torch.uint8, device=dev)
        calls.mae_augment_u8(p(big), p(torch.zeros(1, 5, dtype=torch.int32, device=dev)), 1, 1, 96, 96, 8, p(out), stream())
