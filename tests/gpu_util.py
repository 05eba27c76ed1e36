"""Helpers for the -m gpu parity tests: thin wrappers that call the C ABI on torch CUDA tensors."""
import os

import torch

from ssrl_vit_mae_jepa_b200._lib import calls

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LOG = os.path.join(ROOT, "gpurun_out", "metrics.log")


def report(name, **kv):
    os.makedirs(os.path.dirname(LOG), exist_ok=True)
    line = name + " " + " ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in kv.items())
    with open(LOG, "a") as f:
        f.write(line + "\n")
    print(line)


def stream():
    return torch.cuda.current_stream().cuda_stream


def p(t):
    return None if t is None else t.data_ptr()


def bf(x):
    return x.to(torch.bfloat16).contiguous()


def rel_err(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def max_err(a, b):
    return (a.float() - b.float()).abs().max().item()


def cos_sim(a, b):
    a, b = a.float().flatten().double(), b.float().flatten().double()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def gemm_tn(A, B, epi, bias=None, aux=None):
    M, K = A.shape
    N = B.shape[0]
    out_dtype = torch.float32 if epi == 2 else torch.bfloat16
    out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    pre = torch.empty(M, N, dtype=torch.bfloat16, device=A.device) if epi == 1 else None
    calls.mae_gemm_tn(p(A), p(B), M, N, K, epi, p(bias), p(out), p(pre), p(aux), stream())
    return out, pre


def gemm_wgrad(P, Q, out, so_p, so_q, colsum=None, colsum_of_q=0):
    calls.mae_gemm_wgrad(p(P), P.shape[1], p(Q), Q.shape[1], P.shape[0], p(out), so_p, so_q, p(colsum), colsum_of_q,
                         stream())
