"""Run one mae_gemm_tn shape a few times (for ncu): python tools/gemm_one.py M N K epi [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.gpu_util import gemm_tn, bf
M, N, K, epi = [int(v) for v in sys.argv[1:5]]
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 3
A = bf(torch.randn(M, K, device="cuda")); B = bf(torch.randn(N, K, device="cuda") / K ** 0.5)
bias = torch.randn(N, device="cuda")
aux = torch.randn(M, N, device="cuda") if epi == 2 else (bf(torch.randn(M, N, device="cuda")) if epi == 4 else None)
for _ in range(iters): gemm_tn(A, B, epi, bias, aux)
torch.cuda.synchronize()
print("ok")
