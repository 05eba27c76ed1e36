"""clock64 timeline of CTA 0's second tile in mae_mlp_fwd (debug hook MAE_MLP_TIMELINE): where one tile's time goes."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rows, D = int(sys.argv[1]) if len(sys.argv) > 1 else 290000, int(sys.argv[2]) if len(sys.argv) > 2 else 192
dev = "cuda"
dbg = torch.zeros(3 * 64, dtype=torch.int64, device=dev)
os.environ["MAE_MLP_TIMELINE"] = str(dbg.data_ptr())
from ssrl_vit_mae_jepa_b200._lib import calls  # noqa: E402
from tests.gpu_util import bf, p, stream  # noqa: E402

g = torch.Generator().manual_seed(0)
gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
W1, b1 = bf(torch.randn(4 * D, D, generator=g) / math.sqrt(D)).to(dev), torch.zeros(4 * D, device=dev)
W2, b2 = bf(torch.randn(D, 4 * D, generator=g) / math.sqrt(4 * D)).to(dev), torch.zeros(D, device=dev)
x = torch.randn(rows, D, device=dev)
xn = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
h = torch.empty(rows, 4 * D, dtype=torch.bfloat16, device=dev)
dact = torch.empty(rows, 4 * D, dtype=torch.bfloat16, device=dev)
y = torch.empty(rows, D, device=dev)
for _ in range(3):
    calls.mae_mlp_fwd(p(x), None, 0, 0, rows, D, p(gamma), p(beta), 1e-6, p(W1), p(b1), p(W2), p(b2), p(xn), p(mean),
                      p(rstd), p(h), p(dact), p(y), stream())
torch.cuda.synchronize()
t = dbg.cpu().view(3, 64)
t0 = int(t[0, 0])
names = {0: "epilogue warp 2 (even chunks)", 1: "epilogue warp 10 (odd chunks)", 2: "MMA issuer"}
for r in range(3):
    ev = [(i, int(t[r, i]) - t0) for i in range(64) if int(t[r, i]) != 0]
    print(names[r], " ".join(f"{i}:{v}" for i, v in ev))
