"""Run the encoder-shaped attention (B=2000, L=36, H=6, hd=24: the mma.sync kernels) forward and backward once each after
a warm-up, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssrl_vit_mae_jepa_b200._lib import calls
B, L, H, hd = 2000, 36, 6, 24
st = torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B, L, 3, H, hd, device="cuda").to(torch.bfloat16)
out = torch.empty(B, L, H * hd, dtype=torch.bfloat16, device="cuda"); lse = torch.empty(B, H, L, device="cuda")
dout = torch.randn(B, L, H * hd, device="cuda").to(torch.bfloat16); dqkv = torch.empty_like(qkv)
for _ in range(2):
    calls.mae_attention_fwd(qkv.data_ptr(), B, L, H, hd, out.data_ptr(), lse.data_ptr(), st)
    calls.mae_attention_bwd(qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), B, L, H, hd, dqkv.data_ptr(), st)
torch.cuda.synchronize(); print("ok")
