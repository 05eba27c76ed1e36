"""Golden vectors for the GPU augmentation kernel, produced by the third-party code the reference's transform calls
(torchvision 0.26 ``Compose`` over Pillow 12 images; reference ``src/data.py:15-25``) under ``torch.manual_seed(73)``.

Run from the repo root:  ``python tests/golden/make_augment_golden.py``  ->  ``tests/golden/augment_seed73.npz``
(8 STL-10-shaped uint8 images, the per-sample (i, j, h, w, flip) the transforms drew, and their float32 outputs).
The oracle (``oracle/augment_oracle.py``) must reproduce both the parameters and the pixels exactly before the file
is written.
"""
import os
import sys

import numpy as np
import torch
from PIL import Image
from torchvision import transforms as T

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import augment_oracle as A  # noqa: E402

SEED, N = 73, 8


class Recording(T.RandomResizedCrop):
    """RandomResizedCrop that remembers the parameters it drew (same RNG consumption as the parent)."""

    def forward(self, img):
        i, j, h, w = self.get_params(img, self.scale, self.ratio)
        self.last = (i, j, h, w)
        return T.functional.resized_crop(img, i, j, h, w, self.size, self.interpolation, antialias=self.antialias)


def main():
    rng = np.random.default_rng(SEED)
    base = rng.integers(0, 256, (N, 13, 13, 3)).astype(np.float32)
    imgs = np.clip(np.kron(base, np.ones((1, 8, 8, 1), np.float32))[:, :96, :96] + rng.normal(0, 25, (N, 96, 96, 3)), 0, 255)
    imgs = imgs.astype(np.uint8)
    rrc, flip = Recording(96, scale=(0.8, 1.0)), T.RandomHorizontalFlip()
    post = T.Compose([T.ToTensor(), T.Normalize((0.5, 0.5, 0.5), (0.5, 0.5, 0.5))])
    torch.manual_seed(SEED)
    params, outs = [], []
    for k in range(N):
        im = rrc(Image.fromarray(imgs[k]))
        flipped = torch.rand(1) < flip.p  # RandomHorizontalFlip.forward draws exactly this
        if flipped:
            im = T.functional.hflip(im)
        params.append((*rrc.last, int(flipped)))
        outs.append(post(im).numpy())
    params, expected = np.array(params, np.int32), np.stack(outs)
    nchw = np.ascontiguousarray(imgs.transpose(0, 3, 1, 2))
    torch.manual_seed(SEED)
    assert np.array_equal(A.train_params(N), params), "oracle get_params diverges from torchvision"
    assert np.array_equal(A.augment_batch(nchw, params), expected), "oracle pixels diverge from torchvision/Pillow"
    out = os.path.join(ROOT, "tests", "golden", "augment_seed73.npz")
    np.savez_compressed(out, images_u8=nchw, params=params, expected=expected, seed=np.int64(SEED))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
