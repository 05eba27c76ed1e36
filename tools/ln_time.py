"""Device-timed LayerNorm forward / backward at the shapes of one configs/mae.yaml step (B=2000): decoder rows
(290000 x 192), encoder rows (72000 x 144) and the gathered decoder norm (218000 of 290000 rows)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ssrl_vit_mae_jepa_b200._lib import calls
def t(fn, n=40):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
def run(B, L, D, n_sel=0):
    rows = B * (n_sel or L)
    st = torch.cuda.current_stream().cuda_stream
    xs = [torch.randn(B * L, D, device="cuda") for _ in range(3)]
    gs = [torch.randn(rows, D, device="cuda").to(torch.bfloat16) for _ in range(3)]
    dins = [torch.randn(B * L, D, device="cuda") for _ in range(3)]
    gam, bet = torch.randn(D, device="cuda"), torch.randn(D, device="cuda")
    xn = torch.empty(rows, D, dtype=torch.bfloat16, device="cuda")
    mean, rstd = torch.empty(rows, device="cuda"), torch.empty(rows, device="cuda")
    dx = torch.empty(B * L, D, device="cuda"); dxb = torch.empty(B * L, D, dtype=torch.bfloat16, device="cuda")
    dg, db = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    idx = None
    if n_sel:
        idx = torch.stack([torch.randperm(L - 1, device="cuda")[:n_sel].sort().values + 1 for _ in range(B)]).contiguous()
    ip = idx.data_ptr() if n_sel else None
    f = t(lambda i: calls.mae_layernorm_fwd(xs[i % 3].data_ptr(), gam.data_ptr(), bet.data_ptr(), ip, n_sel, L if n_sel else 0, rows, D,
                                            1e-6, xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(), st))
    b = t(lambda i: calls.mae_layernorm_bwd(gs[i % 3].data_ptr(), xs[i % 3].data_ptr(), mean.data_ptr(), rstd.data_ptr(), gam.data_ptr(),
                                            None if n_sel else dins[i % 3].data_ptr(), ip, n_sel, L if n_sel else 0, rows, D,
                                            dx.data_ptr(), dxb.data_ptr(), dg.data_ptr(), db.data_ptr(), st))
    fb = rows * D * 6.0 + rows * 8
    bb = rows * D * (2 + 4 + (0 if n_sel else 4) + 4 + 2) + rows * 8
    print(f"rows={rows} D={D} gathered={bool(n_sel)} bps={os.environ.get('MAE_LN_FWD_BPS','-')}/{os.environ.get('MAE_LN_BWD_BPS','-')}: "
          f"fwd {f:.1f} us ({fb / f / 1e3:.0f} GB/s)  bwd {b:.1f} us ({bb / b / 1e3:.0f} GB/s)")
run(2000, 145, 192)
run(2000, 36, 144)
run(2000, 145, 192, 109)
