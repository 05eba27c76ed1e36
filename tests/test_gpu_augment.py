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


@pytest.mark.parametrize("H,W,scale,ratio", [(96, 96, (0.8, 1.0), (3 / 4, 4 / 3)), (80, 120, (0.08, 1.0), (3 / 4, 4 / 3)),
                                             (64, 64, (0.9, 1.0), (2.0, 3.0))])
def test_device_drawn_params_match_the_oracle(H, W, scale, ratio):
    """mae_augment_params (boxes + flips drawn on the device, counter-based generator) against the numpy restatement of
    torchvision's RandomResizedCrop.get_params fed with the same uniforms.  Integer outputs: equal, except where CUDA's
    fp32 exp and numpy's differ in the last bit right at a rounding boundary (bounded: < 0.5 % of the samples, by 1 px)."""
    B, seed, counter = 4096, 73, 5
    pr = torch.empty(B, 5, dtype=torch.int32, device=dev)
    calls.mae_augment_params(p(pr), B, H, W, scale[0], scale[1], ratio[0], ratio[1], seed, counter, 1, stream())
    got = pr.cpu().numpy()
    u = A.device_uniforms(B, seed, counter)
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    ref = np.array([A.get_params_from_uniforms(H, W, scale, ratio, u[b]) for b in range(B)], np.int32)
    diff = np.abs(got - ref)
    assert diff.max() <= 1 and (diff.max(1) > 0).mean() < 0.005, (diff.max(), (diff.max(1) > 0).mean())
    i, j, h, w, f = [got[:, k] for k in range(5)]
    assert (h >= 1).all() and (w >= 1).all() and (i >= 0).all() and (j >= 0).all() and (i + h <= H).all() and (j + w <= W).all()
    assert set(np.unique(f)) <= {0, 1} and 0.45 < f.mean() < 0.55
    # a pure function of (seed, counter, sample): repeatable, and a new counter gives new boxes
    pr2 = torch.empty_like(pr)
    calls.mae_augment_params(p(pr2), B, H, W, scale[0], scale[1], ratio[0], ratio[1], seed, counter, 1, stream())
    assert torch.equal(pr, pr2)
    calls.mae_augment_params(p(pr2), B, H, W, scale[0], scale[1], ratio[0], ratio[1], seed, counter + 1, 1, stream())
    assert (pr2 != pr).any(1).float().mean().item() > (0.9 if ratio[0] < 1 else 0.4)  # (fallback boxes are fixed: only the flips change)
    calls.mae_augment_params(p(pr2), B, H, W, scale[0], scale[1], ratio[0], ratio[1], seed, counter, 0, stream())
    assert torch.equal(pr2.cpu(), torch.tensor([[0, 0, H, W, 0]], dtype=torch.int32).expand(B, 5))


def test_gpu_augment_default_draws_on_the_device():
    x = torch.randint(0, 256, (64, 3, 96, 96), dtype=torch.uint8, device=dev)
    aug = GpuAugment(train=True, seed=9)
    a = aug(x)
    pr_a = aug._params.clone()
    b = aug(x)
    assert aug.counter == 2 and not torch.equal(a, b) and aug._next is None  # no host-side draw happened
    ref = A.augment_batch(x[:4].cpu().numpy(), pr_a[:4].cpu().numpy())
    assert np.array_equal(a[:4].cpu().numpy(), ref)
    again = GpuAugment(train=True, seed=9)
    assert torch.equal(again(x), a)  # same seed, same call index -> same augmentation
