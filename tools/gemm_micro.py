"""Micro-benchmark of mae_gemm_tn on given shapes (device-timed, L2 flushed by rotating buffers)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.gpu_util import gemm_tn, bf
shapes = [(290000, 576, 192, 0), (290000, 192, 768, 0), (290000, 192, 192, 0), (290000, 768, 192, 1), (290000, 192, 768, 2), (72000, 432, 144, 0),
          (290000, 768, 192, 4), (290000, 192, 192, 2), (72000, 576, 144, 1), (72000, 576, 144, 4), (72000, 144, 576, 2)]
if len(sys.argv) > 1: shapes = shapes[:int(sys.argv[1])]
for M, N, K, epi in shapes:
    As = [bf(torch.randn(M, K, device="cuda")) for _ in range(3)]
    B = bf(torch.randn(N, K, device="cuda") / K ** 0.5)
    bias = torch.randn(N, device="cuda")
    aux = torch.randn(M, N, device="cuda") if epi == 2 else (bf(torch.randn(M, N, device="cuda")) if epi == 4 else None)
    for i in range(3): gemm_tn(As[i % 3], B, epi, bias, aux)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 12
    for i in range(n): gemm_tn(As[i % 3], B, epi, bias, aux)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    byt = M * K * 2 + M * N * (8 if epi == 2 else (4 if epi in (1, 4) else 2))
    print(f"dbg={os.environ.get('MAE_GEMM_DBG','0')} M={M} N={N} K={K} epi={epi}: {us:7.1f} us  {byt/us/1e6:5.2f} TB/s  {2*M*N*K/us/1e6:6.0f} TFLOP/s")
