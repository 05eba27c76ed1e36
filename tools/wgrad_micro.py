"""Micro-benchmark of mae_gemm_wgrad on the step's shapes (device-timed).  MAE_WGRAD_PCC=1|2 forces the P chunks per CTA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.gpu_util import gemm_wgrad, bf
shapes = [(290000, 768, 192), (290000, 576, 192), (290000, 192, 192), (218000, 192, 192), (72000, 576, 144), (72000, 432, 144), (72000, 144, 144)]
for M, Pc, Qc in shapes:
    Ps = [bf(torch.randn(M, Pc, device="cuda")) for _ in range(2)]
    Q = bf(torch.randn(M, Qc, device="cuda"))
    out = torch.zeros(Pc, Qc, device="cuda"); cs = torch.zeros(Pc, device="cuda")
    for i in range(3): gemm_wgrad(Ps[i % 2], Q, out, Qc, 1, cs, 0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = 12
    for i in range(n): gemm_wgrad(Ps[i % 2], Q, out, Qc, 1, cs, 0)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / n * 1e3
    print(f"pcc={os.environ.get('MAE_WGRAD_PCC','auto')} M={M} Pc={Pc} Qc={Qc}: {us:7.1f} us  {(M*(Pc+Qc)*2)/us/1e6:5.2f} TB/s  {2*M*Pc*Qc/us/1e6:6.0f} TFLOP/s")
