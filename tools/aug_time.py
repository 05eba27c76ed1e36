"""Device-timed cost of the GPU augmentation at BASELINE batch (2000 x 3 x 96 x 96 uint8 -> float32)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssrl_vit_mae_jepa_b200.data import GpuAugment
x = torch.randint(0, 256, (2000, 3, 96, 96), dtype=torch.uint8, device="cuda")
a = GpuAugment(train=True)
pr = a.draw_params(2000, 96, 96, "cuda")
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
us = t(lambda: a(x, pr))
print(f"mae_augment_u8: {us:.1f} us  ({(x.numel() + 4 * x.numel()) / us / 1e6:.2f} TB/s of 1 B in + 4 B out per element)")
print(f"draw_params (torch ops, host-bound): {t(lambda: a.draw_params(2000, 96, 96, 'cuda')):.1f} us")
