"""CPU: pin ``oracle/augment_oracle.py`` against the third-party code the reference's transform calls (Pillow's
bilinear resample, torchvision's RandomResizedCrop / flip / ToTensor / Normalize), bit for bit."""
import os

import numpy as np
import pytest
import torch

from oracle import augment_oracle as A

PIL = pytest.importorskip("PIL.Image")
tvt = pytest.importorskip("torchvision.transforms")

GOLD = os.path.join(os.path.dirname(__file__), "golden", "augment_seed73.npz")


def _img(seed, h=96, w=96):
    rng = np.random.default_rng(seed)
    base = rng.integers(0, 256, (h // 8 + 1, w // 8 + 1, 3)).astype(np.float32)
    up = np.kron(base, np.ones((8, 8, 1), np.float32))[:h, :w]  # blocky structure + noise: edges and flat areas
    return np.clip(up + rng.normal(0, 20, (h, w, 3)), 0, 255).astype(np.uint8)


@pytest.mark.parametrize("box", [(0, 0, 96, 96), (3, 7, 74, 89), (10, 0, 86, 96), (0, 11, 96, 80), (20, 21, 75, 75),
                                 (5, 5, 33, 47), (0, 0, 1, 1), (95, 95, 1, 1), (40, 2, 17, 94)])
def test_resize_matches_pillow(box):
    i, j, h, w = box
    img = _img(h * 100 + w)
    ref = np.asarray(PIL.fromarray(img).crop((j, i, j + w, i + h)).resize((96, 96), PIL.BILINEAR))
    got = A.resize_bilinear_u8(img[i:i + h, j:j + w], 96, 96)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("hw,out", [((200, 150), (96, 96)), ((120, 80), (64, 64)), ((97, 191), (96, 96)), ((96, 96), (32, 48))])
def test_downscale_matches_pillow(hw, out):
    """crops larger than the output use the widened (anti-aliasing) support: more than 2 taps per output pixel"""
    img = _img(hw[0] + hw[1], hw[0], hw[1])
    ref = np.asarray(PIL.fromarray(img).resize((out[1], out[0]), PIL.BILINEAR))
    assert np.array_equal(A.resize_bilinear_u8(img, out[0], out[1]), ref)


def test_coeffs_known_answers():
    # identity: every output pixel reads exactly one input pixel with weight 1 << 22
    xmin, cnt, kk = A.resample_coeffs(96, 96)
    assert np.array_equal(xmin, np.arange(96)) and (kk[:, 0] == 1 << 22).all() and (kk[:, 1:] == 0).all()
    # 1 -> 96: constant image
    assert (A.resize_bilinear_u8(np.full((1, 1, 3), 77, np.uint8), 96, 96) == 77).all()
    # 2x upsampling of a two-pixel ramp: hand-computed triangle weights (centres at .25, .75, 1.25, 1.75)
    xmin, cnt, kk = A.resample_coeffs(2, 4)
    one = 1 << 22
    assert kk[0, :2].tolist() == [one, 0] and kk[1, :2].tolist() == [int(0.75 * one + 0.5), int(0.25 * one + 0.5)]
    assert kk[2, :2].tolist() == [int(0.25 * one + 0.5), int(0.75 * one + 0.5)]
    assert xmin.tolist() == [0, 0, 0, 1] and cnt.tolist() == [1, 2, 2, 1] and kk[3, 0] == one


def test_vectorized_param_draws_follow_get_params():
    """``GpuAugment.draw_params`` (all samples at once, the product's RNG) has the distribution of torchvision's
    sequential ``get_params`` (restated in the oracle): same support, matching moments."""
    from ssrl_vit_mae_jepa_b200.data import GpuAugment

    n = 4000
    g = torch.Generator().manual_seed(1)
    ours = GpuAugment(train=True, generator=g).draw_params(n, 96, 96, "cpu").numpy().astype(np.float64)
    torch.manual_seed(2)
    ref = A.train_params(n).astype(np.float64)
    for col, tol in ((2, 0.5), (3, 0.5), (0, 0.5), (1, 0.5), (4, 0.04)):
        assert abs(ours[:, col].mean() - ref[:, col].mean()) < tol, (col, ours[:, col].mean(), ref[:, col].mean())
    assert abs(ours[:, 2].std() - ref[:, 2].std()) < 0.4 and abs(ours[:, 3].std() - ref[:, 3].std()) < 0.4
    assert ours[:, 2].min() >= 74 and ours[:, 3].min() >= 74 and ours[:, 2].max() <= 96 and ours[:, 3].max() <= 96
    assert ((ours[:, 0] + ours[:, 2]) <= 96).all() and ((ours[:, 1] + ours[:, 3]) <= 96).all() and (ours[:, :2] >= 0).all()
    ev = GpuAugment(train=False).draw_params(3, 96, 96, "cpu")
    assert ev.tolist() == [[0, 0, 96, 96, 0]] * 3


def test_full_transform_matches_torchvision():
    """Same seed, samples processed in order: the oracle's (params, pixels) equal torchvision's Compose exactly."""
    tf = tvt.Compose([tvt.RandomResizedCrop(96, scale=(0.8, 1.0)), tvt.RandomHorizontalFlip(), tvt.ToTensor(),
                      tvt.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    n = 12
    imgs = np.stack([_img(1000 + k) for k in range(n)])  # (n, 96, 96, 3)
    torch.manual_seed(73)
    ref = torch.stack([tf(PIL.fromarray(imgs[k])) for k in range(n)]).numpy()
    torch.manual_seed(73)
    params = A.train_params(n)
    got = A.augment_batch(np.ascontiguousarray(imgs.transpose(0, 3, 1, 2)), params)
    assert np.array_equal(got, ref)
    assert params[:, 4].min() == 0 and params[:, 4].max() == 1  # both flip outcomes occur
    assert (params[:, 2] >= 74).all() and (params[:, 3] >= 74).all()


def test_eval_transform_is_normalize_only():
    ev = tvt.Compose([tvt.Resize(96), tvt.CenterCrop(96), tvt.ToTensor(), tvt.Normalize((0.5,) * 3, (0.5,) * 3)])
    img = _img(5)
    ref = ev(PIL.fromarray(img)).numpy()
    got = A.augment_batch(img.transpose(2, 0, 1)[None], np.array([[0, 0, 96, 96, 0]], np.int32))[0]
    assert np.array_equal(got, ref)


def test_golden_fixture_matches_oracle():
    g = np.load(GOLD)
    got = A.augment_batch(g["images_u8"], g["params"])
    assert np.array_equal(got, g["expected"])
    torch.manual_seed(int(g["seed"]))
    assert np.array_equal(A.train_params(len(g["params"])), g["params"])
