"""clock64 timeline of CTA 0 in the short-row attention forward (debug hook MAE_ENC_TIMELINE): MMA-issuer steps and the
softmax groups' phases for the first items, in cycles relative to the first stamp."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
tl = torch.zeros(4 * 128, dtype=torch.int64, device="cuda")
os.environ["MAE_ENC_TIMELINE"] = str(tl.data_ptr())
from ssrl_vit_mae_jepa_b200._lib import calls
B, L, H, hd = 2000, 36, 6, 24
st = torch.cuda.current_stream().cuda_stream
qkv = torch.randn(B, L, 3, H, hd, device="cuda").to(torch.bfloat16)
out = torch.empty(B, L, H * hd, dtype=torch.bfloat16, device="cuda"); lse = torch.empty(B, H, L, device="cuda")
for _ in range(3):
    calls.mae_attention_fwd(qkv.data_ptr(), B, L, H, hd, out.data_ptr(), lse.data_ptr(), st)
torch.cuda.synchronize()
t = tl.cpu().view(4, 128)
t0 = int(t[t > 0].min())
print("kernel: [entry, setup done, pdl_wait done, thread 0 loop done, all warps done]", [int(x) - t0 if x > 0 else None for x in t[3, 0:5]], "items per CTA ~", 4002 / 148)
last = [int(x) - t0 for x in t[0] if x > 0]
print("last MMA stamp", max(last))
print("MMA issuer, per step: [kv_full ok, s_free ok, S committed, p_ready ok, PV committed]")
for n in range(14):
    print(n, [int(x) - t0 if x > 0 else None for x in t[0, 5 * n:5 * n + 5]])
for g in range(3):
    print(f"softmax group {g}, per item: [s_full ok, pulled, epilogue done, math done, p_free ok, p_ready arrived]")
    for k in range(6):
        print(k, [int(x) - t0 if x > 0 else None for x in t[1 + g, 6 * k:6 * k + 6]])
