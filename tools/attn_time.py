"""Device-timed attention forward / backward at the decoder shape of configs/mae.yaml (B=2000, L=145, H=6, hd=32) and
the encoder shape (L=36, hd=24).  MAE_ATTN_NO_TC=1 selects the mma.sync kernels for the decoder shape."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssrl_vit_mae_jepa_b200._lib import calls
def run(B, L, H, hd):
    D = H * hd
    qkvs = [torch.randn(B, L, 3, H, hd, device="cuda").to(torch.bfloat16) for _ in range(3)]
    out = torch.empty(B, L, D, dtype=torch.bfloat16, device="cuda"); lse = torch.empty(B, H, L, device="cuda")
    dout = torch.randn(B, L, D, device="cuda").to(torch.bfloat16); dqkv = torch.empty_like(qkvs[0])
    st = torch.cuda.current_stream().cuda_stream
    def t(fn, n=12):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    f = t(lambda i: calls.mae_attention_fwd(qkvs[i % 3].data_ptr(), B, L, H, hd, out.data_ptr(), lse.data_ptr(), st))
    b = t(lambda i: calls.mae_attention_bwd(qkvs[i % 3].data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), B, L, H, hd, dqkv.data_ptr(), st))
    ws = torch.empty(B * L * H, device="cuda")
    bw = t(lambda i: calls.mae_attention_bwd_ws(qkvs[i % 3].data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), B, L, H, hd, dqkv.data_ptr(), ws.data_ptr(), ws.numel() * 4, st))
    print(f"   with delta pre-pass (mae_attention_bwd_ws): bwd {bw:.1f} us")
    print(f"B={B} L={L} H={H} hd={hd} no_tc={os.environ.get('MAE_ATTN_NO_TC', '0')}: fwd {f:.1f} us  bwd {b:.1f} us")
run(2000, 145, 6, 32)
run(2000, 36, 6, 24)
def run_q(B, L, H, hd, q_rows):
    D = H * hd
    qkvs = [torch.randn(B, L, 3, H, hd, device="cuda").to(torch.bfloat16) for _ in range(3)]
    out = torch.zeros(B, L, D, dtype=torch.bfloat16, device="cuda"); lse = torch.zeros(B, H, L, device="cuda")
    dout = torch.randn(B, L, D, device="cuda").to(torch.bfloat16); dqkv = torch.empty_like(qkvs[0])
    st = torch.cuda.current_stream().cuda_stream
    def t(fn, n=12):
        for i in range(3): fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n): fn(i)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    f = t(lambda i: calls.mae_attention_fwd_q(qkvs[i % 3].data_ptr(), B, L, H, hd, q_rows, out.data_ptr(), lse.data_ptr(), st))
    b = t(lambda i: calls.mae_attention_bwd_q(qkvs[i % 3].data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), B, L, H, hd, q_rows, dqkv.data_ptr(), st))
    print(f"B={B} L={L} H={H} hd={hd} q_rows={q_rows}: fwd {f:.1f} us  bwd {b:.1f} us")
run_q(2000, 145, 6, 32, 109)
