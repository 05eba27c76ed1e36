"""Micro-benchmark: mae_mlp_fwd (one launch) vs mae_layernorm_fwd + mae_gemm_tn(GELU) + mae_gemm_tn(RESID) at the
decoder / encoder shapes of configs/mae.yaml, B=2000.  CUDA events, L2 flushed by rotating 3 input sets."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ssrl_vit_mae_jepa_b200._lib import calls  # noqa: E402
from tests.gpu_util import bf, gemm_tn, p, stream  # noqa: E402

dev = "cuda"


def run(rows, D, reps=10):
    g = torch.Generator().manual_seed(0)
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    W1, b1 = bf(torch.randn(4 * D, D, generator=g) / math.sqrt(D)).to(dev), torch.zeros(4 * D, device=dev)
    W2, b2 = bf(torch.randn(D, 4 * D, generator=g) / math.sqrt(4 * D)).to(dev), torch.zeros(D, device=dev)
    xs = [torch.randn(rows, D, device=dev) for _ in range(3)]
    xn = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    h = torch.empty(rows, 4 * D, dtype=torch.bfloat16, device=dev)
    dact = torch.empty(rows, 4 * D, dtype=torch.bfloat16, device=dev)
    y = torch.empty(rows, D, device=dev)
    st = stream()

    def fused(x):
        calls.mae_mlp_fwd(p(x), None, 0, 0, rows, D, p(gamma), p(beta), 1e-6, p(W1), p(b1), p(W2), p(b2), p(xn), p(mean),
                          p(rstd), p(h), p(dact), p(y), st)

    def three(x):
        calls.mae_layernorm_fwd(p(x), p(gamma), p(beta), None, 0, 0, rows, D, 1e-6, p(xn), p(mean), p(rstd), st)
        calls.mae_gemm_tn(p(xn), p(W1), rows, 4 * D, D, 1, p(b1), p(h), p(dact), None, st)
        calls.mae_gemm_tn(p(h), p(W2), rows, D, 4 * D, 2, p(b2), p(y), None, p(x), st)

    out = {}
    for name, fn in (("fused", fused), ("three", three)):
        for i in range(3):
            fn(xs[i % 3])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(xs[i % 3])
        e1.record()
        torch.cuda.synchronize()
        out[name] = e0.elapsed_time(e1) * 1e3 / reps
    nbytes = rows * D * 10.0 + rows * 8 + rows * 4 * D * 4.0
    print(f"rows={rows} D={D}: fused {out['fused']:.1f} us ({nbytes / out['fused'] / 1e3:.0f} GB/s), three launches {out['three']:.1f} us")


if __name__ == "__main__":
    run(290000, 192)
    run(218000, 192)
    run(72000, 144)
