"""-m gpu: every kernel of libmae_b200.so, called through the C ABI, against the oracle (integer paths,
bit-exact) or a plain fp32 torch computation of the same op on the same inputs (floating point; tolerances
written at each assert).  bf16-output kernels are compared after rounding the reference to bf16 where that is
the op's contract."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mae_oracle as O
from ssrl_vit_mae_jepa_b200._lib import calls
from tests.gpu_util import bf, cos_sim, gemm_tn, gemm_wgrad, max_err, p, rel_err, report, stream

pytestmark = pytest.mark.gpu
dev = "cuda"


def _mask(noise, L, Lk):
    B = noise.shape[0]
    k = torch.empty(B, Lk, dtype=torch.int64, device=dev)
    m = torch.empty(B, L - Lk, dtype=torch.int64, device=dev)
    calls.mae_random_token_mask(p(noise), B, L, Lk, p(k), p(m), stream())
    return k, m


@pytest.mark.parametrize("B,L,ratio", [(4, 145, 0.75), (257, 145, 0.75), (33, 197, 0.75), (5, 17, 0.6), (3, 256, 0.5),
                                       (2, 145, 0.5), (1, 145, 0.85)])
def test_token_mask_bit_exact(B, L, ratio):
    g = torch.Generator().manual_seed(B * 1000 + L)
    noise = torch.rand(B, L, generator=g)
    Lk = O.num_keep_tokens(L, ratio)
    k, m = _mask(noise.to(dev), L, Lk)
    rk, rm = O.random_token_mask_numpy(noise.numpy(), ratio)
    assert np.array_equal(k.cpu().numpy(), rk) and np.array_equal(m.cpu().numpy(), rm)


def test_token_mask_golden_and_ties(golden):
    k, m = _mask(torch.from_numpy(golden["noise"]).to(dev), 145, 36)
    assert np.array_equal(k.cpu().numpy(), golden["idx_keep"]) and np.array_equal(m.cpu().numpy(), golden["idx_mask"])
    k, m = _mask(torch.from_numpy(golden["tie_noise"]).to(dev), 145, 36)
    assert np.array_equal(k.cpu().numpy(), golden["tie_idx_keep"]) and np.array_equal(m.cpu().numpy(), golden["tie_idx_mask"])
    # heavy ties: noise quantised to 8 levels, and signed zeros
    g = torch.Generator().manual_seed(5)
    noise = (torch.rand(64, 145, generator=g) * 8).floor() / 8
    noise[:, 7] = -0.0
    k, m = _mask(noise.to(dev), 145, 36)
    rk, rm = O.random_token_mask_numpy(noise.numpy(), 0.75)
    assert np.array_equal(k.cpu().numpy(), rk) and np.array_equal(m.cpu().numpy(), rm)


def test_token_mask_full_size_permutation_property():
    noise = torch.rand(2000, 145, device=dev)
    k, m = _mask(noise, 145, 36)
    allidx = torch.cat([k, m], 1)
    assert (allidx.sort(1).values == torch.arange(145, device=dev)).all() and (k[:, 0] == 0).all()
    n = noise.clone(); n[:, 0] = -1
    s = torch.gather(n, 1, allidx)
    assert (s[:, 1:] >= s[:, :-1]).all()  # sortedness
    assert calls.mae_random_token_mask.__name__  # symbol exists
    with pytest.raises(RuntimeError):
        calls.mae_random_token_mask(p(noise), 4, 300, 10, p(k), p(m), stream())  # L > 256 -> MAE_ERR_SHAPE


def test_gather_patches_and_assemble():
    B, Lk, D = 7, 36, 144
    g = torch.Generator().manual_seed(1)
    images = torch.randn(B, 3, 96, 96, generator=g).to(dev)
    noise = torch.rand(B, 145, generator=g)
    k, _ = O.random_token_mask_from_noise(noise, 0.75)
    k = k.to(dev)
    patches = torch.empty(B * (Lk - 1), 192, dtype=torch.bfloat16, device=dev)
    calls.mae_gather_patches(p(images), p(k), B, 3, 96, 8, Lk, p(patches), stream())
    # conv-weight order (c,py,px): unfold gives exactly that
    unf = F.unfold(images, kernel_size=8, stride=8).transpose(1, 2)  # (B,144,192)
    ref = torch.gather(unf, 1, (k[:, 1:] - 1).unsqueeze(-1).expand(-1, -1, 192)).reshape(-1, 192)
    assert torch.equal(patches, bf(ref))
    tok = bf(torch.randn(B * (Lk - 1), D, generator=g)).to(dev)
    cls, pos = torch.randn(D, generator=g).to(dev), torch.randn(145, D, generator=g).to(dev)
    x = torch.empty(B, Lk, D, device=dev)
    calls.mae_assemble_tokens(p(tok), p(cls), p(pos), p(k), B, Lk, D, p(x), stream())
    full = torch.cat([cls.expand(B, 1, D), tok.float().view(B, Lk - 1, D)], 1) + pos[k]
    assert max_err(x, full) < 1e-6


@pytest.mark.parametrize("B,C,img,patch,Lk,D", [(7, 3, 96, 8, 36, 144), (300, 3, 96, 8, 36, 144), (5, 3, 96, 8, 73, 144),
                                                (9, 3, 32, 4, 17, 64), (4, 1, 64, 16, 9, 128), (2000, 3, 96, 8, 36, 144)])
def test_patch_embed_tokens_fused(B, C, img, patch, Lk, D):
    """mae_patch_embed_tokens (gather -> tcgen05 GEMM -> + bias, + pos, class-token rows in one launch): the im2col rows
    bit for bit against F.unfold, and the fp32 stream bit for bit against the three launches it replaces
    (mae_gather_patches + mae_gemm_tn + mae_assemble_tokens: same bf16 operands, same accumulation, same rounding points)."""
    g = torch.Generator().manual_seed(B + D)
    L = (img // patch) ** 2 + 1
    pd = C * patch * patch
    images = torch.randn(B, C, img, img, generator=g).to(dev)
    k = torch.stack([torch.cat([torch.zeros(1, dtype=torch.int64), 1 + torch.randperm(L - 1, generator=g)[:Lk - 1]])
                     for _ in range(B)]).to(dev)
    W, bias = bf(torch.randn(D, pd, generator=g) / math.sqrt(pd)).to(dev), torch.randn(D, generator=g).to(dev)
    cls, pos = torch.randn(D, generator=g).to(dev), torch.randn(L, D, generator=g).to(dev)
    patches = torch.zeros(B * (Lk - 1), pd, dtype=torch.bfloat16, device=dev)
    x = torch.full((B, Lk, D), float("nan"), device=dev)
    calls.mae_patch_embed_tokens(p(images), p(k), B, C, img, patch, Lk, D, p(W), p(bias), p(cls), p(pos), p(patches), p(x),
                                 stream())
    unf = F.unfold(images, kernel_size=patch, stride=patch).transpose(1, 2)  # (B, L-1, pd) in conv-weight order (c,py,px)
    ref = torch.gather(unf, 1, (k[:, 1:] - 1).unsqueeze(-1).expand(-1, -1, pd)).reshape(-1, pd)
    assert torch.equal(patches, bf(ref))
    p3 = torch.empty_like(patches)
    calls.mae_gather_patches(p(images), p(k), B, C, img, patch, Lk, p(p3), stream())
    tok, _ = gemm_tn(p3, W, 0, bias)
    x3 = torch.empty(B, Lk, D, device=dev)
    calls.mae_assemble_tokens(p(tok), p(cls), p(pos), p(k), B, Lk, D, p(x3), stream())
    assert torch.equal(x, x3), max_err(x, x3)
    full = torch.cat([cls.expand(B, 1, D), bf(ref.float() @ W.float().t() + bias).float().view(B, Lk - 1, D)], 1) + pos[k]
    assert max_err(x, full) <= 2 ** -7 * full.abs().max().item()  # vs fp32 torch: one bf16 rounding of the conv output


@pytest.mark.parametrize("rows,D", [(1000, 144), (777, 192), (5, 16), (300, 256), (500, 512), (401, 768), (97, 1024)])
def test_layernorm_fwd_bwd(rows, D):
    g = torch.Generator().manual_seed(rows)
    x = (torch.randn(rows, D, generator=g) * 2 + 0.5).to(dev)
    gamma, beta = (torch.rand(D, generator=g) + 0.5).to(dev), torch.randn(D, generator=g).to(dev)
    xn = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    calls.mae_layernorm_fwd(p(x), p(gamma), p(beta), None, 0, 0, rows, D, 1e-6, p(xn), p(mean), p(rstd), stream())
    xr = x.clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    ref = F.layer_norm(xr, (D,), gr, br, 1e-6)
    assert max_err(xn, ref) <= 2 ** -7 * ref.abs().max().item() + 1e-6  # bf16 rounding of the output
    assert max_err(mean, x.mean(1)) < 1e-5 and rel_err(rstd, 1 / torch.sqrt(x.var(1, unbiased=False) + 1e-6)) < 1e-5
    dy = bf(torch.randn(rows, D, generator=g)).to(dev)
    dx_in = torch.randn(rows, D, generator=g).to(dev)
    ref.backward(dy.float())
    dx = torch.empty(rows, D, device=dev)
    dxb = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    calls.mae_layernorm_bwd(p(dy), p(x), p(mean), p(rstd), p(gamma), p(dx_in), None, 0, 0, rows, D, p(dx), p(dxb), p(dg), p(db), stream())
    e1, e2, e3 = max_err(dx, xr.grad + dx_in), rel_err(dg, gr.grad), rel_err(db, br.grad)
    report("layernorm", rows=rows, D=D, dx_max=e1, dgamma_rel=e2, dbeta_rel=e3)
    assert e1 < 1e-4 and e2 < 1e-4 and e3 < 1e-4  # fp32 kernel vs fp32 torch
    assert torch.equal(dxb, bf(dx))


def test_layernorm_gather_variants():
    B, L, n_sel, D = 6, 145, 109, 192
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B * L, D, generator=g).to(dev)
    idx = torch.stack([torch.randperm(L, generator=g)[:n_sel] for _ in range(B)]).to(dev)
    gamma, beta = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    rows = B * n_sel
    xn = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    calls.mae_layernorm_fwd(p(x), p(gamma), p(beta), p(idx), n_sel, L, rows, D, 1e-6, p(xn), p(mean), p(rstd), stream())
    sel = torch.gather(x.view(B, L, D), 1, idx.unsqueeze(-1).expand(-1, -1, D)).reshape(rows, D)
    assert max_err(xn, F.layer_norm(sel, (D,), eps=1e-6)) < 0.04
    dy = bf(torch.randn(rows, D, generator=g)).to(dev)
    dx = torch.zeros(B * L, D, device=dev)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    calls.mae_layernorm_bwd(p(dy), p(x), p(mean), p(rstd), p(gamma), None, p(idx), n_sel, L, rows, D, p(dx), None, p(dg), p(db), stream())
    xr = x.clone().view(B, L, D).requires_grad_(True)
    F.layer_norm(torch.gather(xr, 1, idx.unsqueeze(-1).expand(-1, -1, D)), (D,), eps=1e-6).backward(dy.float().view(B, n_sel, D))
    assert max_err(dx.view(B, L, D), xr.grad) < 1e-4


def test_layernorm_masked_tail_variants():
    """mae_layernorm_fwd_add (gathered x + compact bf16 delta -> LN, summed rows kept) and mae_layernorm_bwd_scatter
    (compact residual gradient, output scattered to the mapped rows): the two norms around the masked-rows-only tail of
    the last decoder block."""
    B, L, n_sel, D = 5, 145, 109, 192
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B * L, D, generator=g).to(dev)
    idx = torch.stack([torch.randperm(L, generator=g)[:n_sel].sort().values for _ in range(B)]).to(dev)
    rows = B * n_sel
    delta = bf(torch.randn(rows, D, generator=g)).to(dev)
    gamma, beta = (torch.rand(D, generator=g) + 0.5).to(dev), torch.randn(D, generator=g).to(dev)
    xn = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    x_out = torch.empty(rows, D, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    calls.mae_layernorm_fwd_add(p(x), p(delta), p(gamma), p(beta), p(idx), n_sel, L, rows, D, 1e-6, p(x_out), p(xn), p(mean),
                                p(rstd), stream())
    sel = torch.gather(x.view(B, L, D), 1, idx.unsqueeze(-1).expand(-1, -1, D)).reshape(rows, D) + delta.float()
    assert torch.equal(x_out, sel)
    ref = F.layer_norm(sel, (D,), gamma, beta, 1e-6)
    assert max_err(xn, ref) <= 2 ** -7 * ref.abs().max().item() + 1e-6
    assert max_err(mean, sel.mean(1)) < 1e-5
    # non-gathered flavour (x already compact)
    x2 = torch.randn(rows, D, generator=g).to(dev)
    calls.mae_layernorm_fwd_add(p(x2), p(delta), p(gamma), p(beta), None, 0, 0, rows, D, 1e-6, p(x_out), p(xn), p(mean),
                                p(rstd), stream())
    assert torch.equal(x_out, x2 + delta.float())
    # backward: compact dy and compact residual gradient, x gathered, dx scattered to the mapped rows
    calls.mae_layernorm_fwd(p(x), p(gamma), p(beta), p(idx), n_sel, L, rows, D, 1e-6, p(xn), p(mean), p(rstd), stream())
    dy = bf(torch.randn(rows, D, generator=g)).to(dev)
    din = torch.randn(rows, D, generator=g).to(dev)
    dx = torch.full((B * L, D), 7.0, device=dev)
    dxb = torch.full((B * L, D), 7.0, dtype=torch.bfloat16, device=dev)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    calls.mae_layernorm_bwd_scatter(p(dy), p(x), p(mean), p(rstd), p(gamma), p(din), p(idx), n_sel, L, rows, D, p(dx), p(dxb),
                                    p(dg), p(db), stream())
    xr = x.clone().view(B, L, D).requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    xs = torch.gather(xr, 1, idx.unsqueeze(-1).expand(-1, -1, D))
    F.layer_norm(xs, (D,), gr, br, 1e-6).backward(dy.float().view(B, n_sel, D))
    want = torch.full((B, L, D), 7.0, device=dev)  # rows that are not selected stay untouched
    want.scatter_(1, idx.unsqueeze(-1).expand(-1, -1, D),
                  torch.gather(xr.grad, 1, idx.unsqueeze(-1).expand(-1, -1, D)) + din.view(B, n_sel, D))
    assert max_err(dx.view(B, L, D), want) < 1e-4
    assert torch.equal(dxb, bf(dx))
    assert rel_err(dg, gr.grad) < 1e-4 and rel_err(db, br.grad) < 1e-4


def test_scatter_decoder_tokens_ln_matches_two_launches():
    """mae_scatter_decoder_tokens_ln = mae_scatter_decoder_tokens followed by mae_layernorm_fwd: the stream bit for bit,
    the normalised rows and statistics like the separate LayerNorm launch (up to FMA contraction)."""
    B, L, nk, D = 7, 145, 36, 192
    g = torch.Generator().manual_seed(29)
    perm = torch.stack([torch.cat([torch.zeros(1, dtype=torch.int64), torch.randperm(L - 1, generator=g) + 1]) for _ in range(B)])
    idx_keep, idx_mask = perm[:, :nk].contiguous().to(dev), perm[:, nk:].contiguous().to(dev)
    emb = bf(torch.randn(B * nk, D, generator=g)).to(dev)
    mt, pos = torch.randn(D, generator=g).to(dev), torch.randn(L, D, generator=g).to(dev)
    gamma, beta = (torch.rand(D, generator=g) + 0.5).to(dev), torch.randn(D, generator=g).to(dev)
    x1, x2 = torch.empty(B * L, D, device=dev), torch.empty(B * L, D, device=dev)
    xn1, xn2 = (torch.empty(B * L, D, dtype=torch.bfloat16, device=dev) for _ in range(2))
    m1, r1, m2, r2 = (torch.empty(B * L, device=dev) for _ in range(4))
    calls.mae_scatter_decoder_tokens(p(emb), p(mt), p(pos), p(idx_keep), p(idx_mask), B, L, nk, D, p(x1), stream())
    calls.mae_layernorm_fwd(p(x1), p(gamma), p(beta), None, 0, 0, B * L, D, 1e-6, p(xn1), p(m1), p(r1), stream())
    calls.mae_scatter_decoder_tokens_ln(p(emb), p(mt), p(pos), p(idx_keep), p(idx_mask), B, L, nk, D, p(gamma), p(beta), 1e-6,
                                        p(x2), p(xn2), p(m2), p(r2), stream())
    assert torch.equal(x1, x2)
    # same arithmetic in the same order; only the compiler's FMA contraction may differ between the two kernels
    assert max_err(m1, m2) < 1e-6 and rel_err(r1, r2) < 1e-6
    assert max_err(xn1, xn2) <= 2 ** -7 * xn1.float().abs().max().item() and (xn1 != xn2).float().mean().item() < 1e-3
    ref = torch.empty(B, L, D, device=dev)
    ref.scatter_(1, idx_keep.unsqueeze(-1).expand(-1, -1, D), emb.float().view(B, nk, D))
    ref.scatter_(1, idx_mask.unsqueeze(-1).expand(-1, -1, D), mt.expand(B, L - nk, D).contiguous())
    assert torch.equal(x2.view(B, L, D), ref + pos)


GEMM_SHAPES = [(128, 64, 64), (128, 144, 144), (300, 432, 144), (1000, 576, 192), (777, 144, 576), (517, 192, 768),
               (128, 768, 192), (2500, 192, 192), (64, 144, 192), (40000, 576, 192)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_tn_bias_bf16(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    A, B = bf(torch.randn(M, K, generator=g)).to(dev), bf(torch.randn(N, K, generator=g) / math.sqrt(K)).to(dev)
    bias = torch.randn(N, generator=g).to(dev)
    out, _ = gemm_tn(A, B, 0, bias)
    ref = A.float() @ B.float().t() + bias
    e = max_err(out, ref)
    report("gemm_tn", M=M, N=N, K=K, max_err=e, rel=rel_err(out, ref))
    assert e <= 2 ** -7 * ref.abs().max().item()  # one bf16 ulp of the largest output
    out2, _ = gemm_tn(A, B, 0, None)
    assert max_err(out2, A.float() @ B.float().t()) <= 2 ** -7 * ref.abs().max().item()


def test_gemm_tn_epilogues():
    M, D = 1000, 192
    g = torch.Generator().manual_seed(9)
    x = bf(torch.randn(M, D, generator=g)).to(dev)
    W1, b1 = bf(torch.randn(4 * D, D, generator=g) / math.sqrt(D)).to(dev), torch.randn(4 * D, generator=g).to(dev)
    h, dact = gemm_tn(x, W1, 1, b1)   # h = gelu(pre), dact = gelu'(pre) with pre = bf16(x W1^T + b1)
    pre_ref = bf(x.float() @ W1.float().t() + b1).float().requires_grad_(True)
    h_ref = F.gelu(pre_ref)
    h_ref.sum().backward()
    # bf16 outputs of a 1.5e-5-accurate GELU evaluated on a pre-activation that itself may round differently by 1 ulp
    assert max_err(h, h_ref) <= 2 ** -7 * pre_ref.abs().max().item()
    assert max_err(dact, pre_ref.grad) <= 2 ** -7 * 1.2 + 2e-2  # |gelu'| <= 1.13; its slope is <= 0.4 per unit of pre
    report("gemm_tn_gelu", h_err=max_err(h, h_ref), dact_err=max_err(dact, pre_ref.grad))
    W2, b2 = bf(torch.randn(D, 4 * D, generator=g) / math.sqrt(4 * D)).to(dev), torch.randn(D, generator=g).to(dev)
    resid = torch.randn(M, D, generator=g).to(dev)
    y, _ = gemm_tn(h, W2, 2, b2, aux=resid)
    y_ref = resid + h.float() @ W2.float().t() + b2
    e = max_err(y, y_ref)
    report("gemm_tn_resid", max_err=e)
    assert y.dtype == torch.float32 and e < 2e-4 * max(1.0, y_ref.abs().max().item())  # fp32 out, fp32 accumulate
    dy = bf(torch.randn(M, D, generator=g)).to(dev)
    W2t = bf(W2.float().t())  # (4D, D): B operand of the dgrad
    dh, _ = gemm_tn(dy, W2t, 4, None, aux=dact)
    dh_ref = (dy.float() @ W2.float()) * dact.float()
    e = max_err(dh, dh_ref)
    report("gemm_tn_dgelu", max_err=e)
    assert e <= 2 ** -7 * dh_ref.abs().max().item()
    # and against autograd of exact GELU at the bf16 pre-activation (dact carries one bf16 rounding)
    e2 = max_err(dh, (dy.float() @ W2.float()) * pre_ref.grad)
    assert e2 <= 2 ** -6 * dh_ref.abs().max().item()


def _mlp_fwd(x, idx, n_sel, L_src, rows, D, gamma, beta, W1, b1, W2, b2):
    xn = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
    mean, rstd = torch.empty(rows, device=dev), torch.empty(rows, device=dev)
    h = torch.empty(rows, 4 * D, dtype=torch.bfloat16, device=dev)
    dact = torch.empty(rows, 4 * D, dtype=torch.bfloat16, device=dev)
    y = torch.empty(rows, D, device=dev)
    calls.mae_mlp_fwd(p(x), p(idx), n_sel, L_src, rows, D, p(gamma), p(beta), 1e-6, p(W1), p(b1), p(W2), p(b2), p(xn), p(mean),
                      p(rstd), p(h), p(dact), p(y), stream())
    return xn, mean, rstd, h, dact, y


@pytest.mark.parametrize("rows,D,gather", [(1000, 192, False), (777, 144, False), (128, 64, False), (40000, 192, False),
                                           (20001, 144, False), (5 * 109, 192, True), (300 * 109, 192, True), (2 * 109, 96, True)])
def test_mlp_fwd_fused(rows, D, gather):
    """mae_mlp_fwd (LayerNorm -> fc1 -> GELU -> fc2 -> + residual in one launch) against fp32 torch, step by step on the
    kernel's own intermediate outputs (so each stage is checked to its own rounding), and against the three-launch path
    (mae_layernorm_fwd + mae_gemm_tn EPI_GELU + mae_gemm_tn EPI_RESID) it replaces."""
    g = torch.Generator().manual_seed(rows + D)
    L_src, n_sel = (145, 109) if gather else (0, 0)
    nsrc = (rows // n_sel) * L_src if gather else rows
    x = (torch.randn(nsrc, D, generator=g) * 1.5 + 0.3).to(dev)
    idx = None
    if gather:
        B = rows // n_sel
        idx = torch.stack([torch.randperm(L_src, generator=g)[:n_sel] for _ in range(B)]).to(dev)  # (B, n_sel) int64
        src_rows = (torch.arange(B, device=dev)[:, None] * L_src + idx).reshape(-1)
        xs = x[src_rows]
    else:
        xs = x
    gamma, beta = (1 + 0.1 * torch.randn(D, generator=g)).to(dev), (0.1 * torch.randn(D, generator=g)).to(dev)
    W1, b1 = bf(torch.randn(4 * D, D, generator=g) / math.sqrt(D)).to(dev), (0.5 * torch.randn(4 * D, generator=g)).to(dev)
    W2, b2 = bf(torch.randn(D, 4 * D, generator=g) / math.sqrt(4 * D)).to(dev), torch.randn(D, generator=g).to(dev)
    xn, mean, rstd, h, dact, y = _mlp_fwd(x, idx, n_sel, L_src, rows, D, gamma, beta, W1, b1, W2, b2)
    # LayerNorm (fp32 statistics, bf16 output)
    mu = xs.mean(1)
    rs = (xs.var(1, unbiased=False) + 1e-6).rsqrt()
    assert max_err(mean, mu) < 1e-5 and rel_err(rstd, rs) < 1e-5
    xn_ref = F.layer_norm(xs, (D,), gamma, beta, 1e-6)
    assert max_err(xn, xn_ref) <= 2 ** -8 * xn_ref.abs().max().item() + 1e-6   # one bf16 rounding
    # fc1 + GELU / GELU' on the kernel's own xn (pre = bf16(xn W1^T + b1))
    pre_ref = bf(xn.float() @ W1.float().t() + b1).float().requires_grad_(True)
    h_ref = F.gelu(pre_ref)
    h_ref.sum().backward()
    assert max_err(h, h_ref) <= 2 ** -7 * pre_ref.abs().max().item()
    assert max_err(dact, pre_ref.grad) <= 2 ** -7 * 1.2 + 2e-2
    # fc2 + bias + residual on the kernel's own h (fp32 accumulate, fp32 out)
    y_ref = xs + h.float() @ W2.float().t() + b2
    e = max_err(y, y_ref)
    assert e < 2e-4 * max(1.0, y_ref.abs().max().item())
    # the three launches it replaces
    xn3 = torch.empty_like(xn); m3 = torch.empty_like(mean); r3 = torch.empty_like(rstd)
    calls.mae_layernorm_fwd(p(x), p(gamma), p(beta), p(idx), n_sel, L_src, rows, D, 1e-6, p(xn3), p(m3), p(r3), stream())
    h3, d3 = gemm_tn(xn3, W1, 1, b1)
    y3, _ = gemm_tn(h3, W2, 2, b2, aux=xs.contiguous())
    report("mlp_fwd_fused", rows=rows, D=D, gather=int(gather), y_err=e, xn_vs3=max_err(xn, xn3), h_vs3=rel_err(h, h3),
           y_vs3=rel_err(y, y3))
    assert rel_err(xn, xn3) < 1e-3 and rel_err(h, h3) < 5e-3 and rel_err(dact, d3) < 5e-3 and rel_err(y, y3) < 2e-3


WGRAD_SHAPES = [(64, 128, 64), (1000, 432, 144), (1000, 144, 144), (5000, 768, 192), (70000, 576, 144), (333, 192, 192),
                (218000, 192, 192), (3000, 2304, 768), (1000, 768, 512), (900, 512, 304)]


@pytest.mark.parametrize("M,Pc,Qc", WGRAD_SHAPES)
def test_gemm_wgrad(M, Pc, Qc):
    g = torch.Generator().manual_seed(M + Pc)
    P, Q = bf(torch.randn(M, Pc, generator=g)).to(dev), bf(torch.randn(M, Qc, generator=g)).to(dev)
    out = torch.zeros(Pc, Qc, device=dev)
    gemm_wgrad(P, Q, out, Qc, 1)
    ref = (P.double().t() @ Q.double()).float()
    e = rel_err(out, ref)
    report("gemm_wgrad", M=M, Pc=Pc, Qc=Qc, rel=e, max_err=max_err(out, ref))
    assert e < 1e-5  # exact bf16 products, fp32 accumulation
    outT = torch.ones(Qc, Pc, device=dev)  # accumulate semantics + transposed strides + fused bias gradient (of P)
    cs = torch.zeros(Pc, device=dev)
    gemm_wgrad(P, Q, outT, 1, Pc, cs, 0)
    assert rel_err(outT - 1, ref.t()) < 1e-5 + 1e-6 * M ** 0.5
    assert max_err(cs, P.double().sum(0).float()) < 1e-3 * M ** 0.5
    csq = torch.zeros(Qc, device=dev)  # ... and of Q
    out2 = torch.zeros(Pc, Qc, device=dev)
    gemm_wgrad(P, Q, out2, Qc, 1, csq, 1)
    assert rel_err(out2, ref) < 1e-5 and max_err(csq, Q.double().sum(0).float()) < 1e-3 * M ** 0.5


@pytest.mark.parametrize("B,L,H,hd", [(3, 36, 6, 24), (2, 145, 6, 32), (2, 145, 6, 24), (5, 17, 2, 32), (2, 73, 6, 24),
                                      (1, 160, 1, 32), (1, 1, 1, 24), (2, 49, 12, 64), (2, 197, 16, 32), (1, 197, 12, 64),
                                      (1, 208, 2, 64), (2, 99, 3, 64),
                                      # the tcgen05 forward (128 < L <= 160, hd 32, even H): ragged L, one / many items per CTA
                                      (1, 129, 2, 32), (3, 160, 4, 32), (2, 131, 6, 32), (250, 145, 6, 32),
                                      # the short-row tcgen05 kernels (L <= 64, hd <= 32): every row pitch (16 / 32 / 40 / 48 / 64),
                                      # batches that are not a multiple of the samples per tile, head dims 8 .. 32
                                      (7, 36, 6, 24), (500, 36, 6, 24), (9, 16, 3, 8), (5, 33, 2, 16), (4, 40, 1, 32), (5, 48, 2, 24),
                                      (3, 64, 4, 32), (3, 57, 2, 24), (17, 9, 5, 32)])
def test_attention_fwd_bwd(B, L, H, hd):
    D = H * hd
    g = torch.Generator().manual_seed(L * 7 + hd)
    qkv = bf(torch.randn(B, L, 3, H, hd, generator=g)).to(dev)
    out = torch.empty(B, L, D, dtype=torch.bfloat16, device=dev)
    lse = torch.empty(B, H, L, device=dev)
    calls.mae_attention_fwd(p(qkv), B, L, H, hd, p(out), p(lse), stream())
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.permute(2, 0, 3, 1, 4).unbind(0)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, D)
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    e = max_err(out, ref)
    report("attn_fwd", L=L, hd=hd, max_err=e, lse_err=max_err(lse, torch.logsumexp(s, -1)))
    assert e < 2e-2 and max_err(lse, torch.logsumexp(s, -1)) < 1e-3  # P rounded to bf16, bf16 output
    dout = bf(torch.randn(B, L, D, generator=g)).to(dev)
    ref.backward(dout.float())
    dqkv = torch.empty_like(qkv)
    calls.mae_attention_bwd(p(qkv), p(out), p(dout), p(lse), B, L, H, hd, p(dqkv), stream())
    for i, name in enumerate("qkv"):
        c = cos_sim(dqkv[:, :, i], qf.grad[:, :, i])
        r = rel_err(dqkv[:, :, i], qf.grad[:, :, i])
        report("attn_bwd", L=L, hd=hd, which=name, cos=c, rel=r)
        if L > 1:
            assert c > 0.999 and r < 3e-2  # bf16 P / dS operands
        else:
            assert max_err(dqkv[:, :, i], qf.grad[:, :, i]) < 2e-2


@pytest.mark.parametrize("B,L,H", [(1, 145, 2), (74, 145, 6), (149, 133, 4), (301, 160, 2)])
def test_attention_tc_repeatable_and_item_distribution(B, L, H):
    """The tcgen05 attention kernels are persistent (148 CTAs walking (sample, head) items through mbarrier pipelines) and
    use no atomics: results must be bit-identical from run to run and correct for 1 item, fewer / more items than CTAs and
    item counts that are not multiples of the grid (a lost or reordered pipeline phase shows up as a mismatch)."""
    hd, D = 32, H * 32
    g = torch.Generator().manual_seed(B + L)
    qkv = bf(torch.randn(B, L, 3, H, hd, generator=g)).to(dev)
    dout = bf(torch.randn(B, L, D, generator=g)).to(dev)
    outs = []
    for _ in range(3):
        out = torch.full((B, L, D), float("nan"), dtype=torch.bfloat16, device=dev)
        lse = torch.full((B, H, L), float("nan"), device=dev)
        dqkv = torch.full_like(qkv, float("nan"))
        calls.mae_attention_fwd(p(qkv), B, L, H, hd, p(out), p(lse), stream())
        calls.mae_attention_bwd(p(qkv), p(out), p(dout), p(lse), B, L, H, hd, p(dqkv), stream())
        outs.append((out, lse, dqkv))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
    out, lse, dqkv = outs[0]
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all() and torch.isfinite(dqkv.float()).all()
    # the workspace entry point (delta = rowsum(dO o O) from a coalesced pre-pass instead of per-thread gathers): same sums
    ws = torch.full((B * L * H + 64,), float("nan"), device=dev)
    dqkv_ws = torch.full_like(qkv, float("nan"))
    calls.mae_attention_bwd_ws(p(qkv), p(out), p(dout), p(lse), B, L, H, hd, p(dqkv_ws), p(ws), B * L * H * 4, stream())
    assert torch.isnan(ws[B * L * H:]).all()  # the scratch is not overrun
    ref_delta = (out.float() * dout.float()).view(B, L, H, hd).sum(-1)
    assert max_err(ws[:B * L * H].view(B, L, H), ref_delta) < 1e-4 * hd
    assert rel_err(dqkv_ws, dqkv) < 2e-3
    # spot-check samples spread over the batch (first, middle, last) against fp32 torch
    for b in sorted({0, B // 2, B - 1}):
        qf = qkv[b:b + 1].float().requires_grad_(True)
        q, k, v = qf.permute(2, 0, 3, 1, 4).unbind(0)
        ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(1, L, D)
        ref.backward(dout[b:b + 1].float())
        assert max_err(out[b:b + 1], ref) < 2e-2
        assert rel_err(dqkv[b:b + 1], qf.grad) < 3e-2


@pytest.mark.parametrize("B,L,H,hd", [(1, 36, 1, 24), (2, 36, 6, 24), (148 * 3 + 1, 36, 2, 24), (1000, 36, 6, 24), (301, 17, 2, 32)])
def test_attention_enc_repeatable_and_item_distribution(B, L, H, hd):
    """The short-row tcgen05 attention (encoder shape: several samples per 128-row tile, two softmax groups alternating over a
    persistent CTA's items): bit-identical from run to run, correct for one item, fewer / more items than CTAs, odd item counts
    per CTA (the two groups get different numbers of items) and a last tile with missing samples; outputs outside the tensors'
    live rows must not be touched (canary rows of NaN stay NaN-free inside, and a guard region after the buffers stays intact)."""
    D = H * hd
    g = torch.Generator().manual_seed(B + L)
    qkv = bf(torch.randn(B, L, 3, H, hd, generator=g)).to(dev)
    dout = bf(torch.randn(B, L, D, generator=g)).to(dev)
    outs = []
    for _ in range(3):
        out = torch.full((B + 1, L, D), float("nan"), dtype=torch.bfloat16, device=dev)
        lse = torch.full((B + 1, H, L), float("nan"), device=dev)
        dqkv = torch.full((B + 1, L, 3, H, hd), float("nan"), dtype=torch.bfloat16, device=dev)
        calls.mae_attention_fwd(p(qkv), B, L, H, hd, p(out), p(lse), stream())
        calls.mae_attention_bwd(p(qkv), p(out), p(dout), p(lse), B, L, H, hd, p(dqkv), stream())
        assert torch.isnan(out[B].float()).all() and torch.isnan(lse[B]).all() and torch.isnan(dqkv[B].float()).all()  # guard
        outs.append((out[:B], lse[:B], dqkv[:B]))
    for o in outs[1:]:
        assert torch.equal(o[0], outs[0][0]) and torch.equal(o[1], outs[0][1]) and torch.equal(o[2], outs[0][2])
    out, lse, dqkv = outs[0]
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all() and torch.isfinite(dqkv.float()).all()
    qf = qkv.float().requires_grad_(True)
    q, k, v = qf.permute(2, 0, 3, 1, 4).unbind(0)
    ref = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, L, D)
    ref.backward(dout.float())
    assert max_err(out, ref) < 2e-2
    for b in sorted({0, B // 2, B - 1}):  # per-sample bound as well: a swapped sample slot would hide in a global norm
        assert rel_err(dqkv[b], qf.grad[b]) < 3e-2
    assert rel_err(dqkv, qf.grad) < 3e-2


@pytest.mark.parametrize("B,L,H,q_rows", [(3, 145, 6, 109), (150, 145, 2, 109), (301, 133, 4, 128), (2, 160, 2, 73)])
def test_attention_leading_query_rows(B, L, H, q_rows):
    """mae_attention_fwd_q / bwd_q (only the first q_rows query tokens matter: the decoder-sized tcgen05 kernels drop their
    second query tile): the wanted output rows equal the full kernel's, the gradient equals the full kernel's gradient for a
    dout that is zero on the other rows -- even when the dout passed in is NOT zero there."""
    hd, D = 32, H * 32
    g = torch.Generator().manual_seed(B * 3 + L)
    qkv = bf(torch.randn(B, L, 3, H, hd, generator=g)).to(dev)
    dout = bf(torch.randn(B, L, D, generator=g)).to(dev)
    out = torch.full((B, L, D), float("nan"), dtype=torch.bfloat16, device=dev)
    lse = torch.full((B, H, L), float("nan"), device=dev)
    out_q, lse_q = out.clone(), lse.clone()
    calls.mae_attention_fwd(p(qkv), B, L, H, hd, p(out), p(lse), stream())
    calls.mae_attention_fwd_q(p(qkv), B, L, H, hd, q_rows, p(out_q), p(lse_q), stream())
    assert torch.equal(out_q[:, :q_rows], out[:, :q_rows]) and torch.equal(lse_q[:, :, :q_rows], lse[:, :, :q_rows])
    assert torch.isfinite(out_q.float()).all()  # rows that are not wanted may hold anything finite (they feed GEMMs)
    dout0 = dout.clone(); dout0[:, q_rows:] = 0
    ref = torch.full_like(qkv, float("nan")); got = torch.full_like(qkv, float("nan"))
    calls.mae_attention_bwd(p(qkv), p(out), p(dout0), p(lse), B, L, H, hd, p(ref), stream())
    calls.mae_attention_bwd_q(p(qkv), p(out_q), p(dout), p(lse_q), B, L, H, hd, q_rows, p(got), stream())
    assert torch.isfinite(got.float()).all()
    assert (got[:, q_rows:, 0] == 0).all()  # no query gradient on the other rows
    assert rel_err(got, ref) < 2e-3 and max_err(got, ref) < 2e-2


def test_scatter_decoder_and_row_ops():
    B, L, Lk, D = 5, 145, 36, 192
    g = torch.Generator().manual_seed(11)
    noise = torch.rand(B, L, generator=g)
    k, m = [t.to(dev) for t in O.random_token_mask_from_noise(noise, 0.75)]
    emb = bf(torch.randn(B * Lk, D, generator=g)).to(dev)
    mt, pos = torch.randn(D, generator=g).to(dev), torch.randn(L, D, generator=g).to(dev)
    x = torch.empty(B, L, D, device=dev)
    calls.mae_scatter_decoder_tokens(p(emb), p(mt), p(pos), p(k), p(m), B, L, Lk, D, p(x), stream())
    ref = O.set_at_index(O.repeat_token(mt.view(1, 1, D), (B, L)), k, emb.float().view(B, Lk, D)) + pos
    assert max_err(x, ref) < 1e-6
    # grouped layout: [mask tokens in idx_mask order | kept tokens in idx_keep order] = the same rows, permuted
    xg = torch.full((B, L, D), float("nan"), device=dev)
    calls.mae_scatter_decoder_tokens_grouped(p(emb), p(mt), p(pos), p(k), p(m), B, L, Lk, D, p(xg), stream())
    assert torch.equal(xg, torch.cat([O.get_at_index(x, m), O.get_at_index(x, k)], dim=1))
    src = torch.randn(B, L, D, generator=g).to(dev)
    out = torch.empty(B * Lk, D, dtype=torch.bfloat16, device=dev)
    calls.mae_gather_rows_bf16(p(src), p(k), B, Lk, L, 0, D, p(out), stream())
    assert torch.equal(out, bf(O.get_at_index(src, k).reshape(-1, D)))
    out2 = torch.empty(B * (L - 1), D, dtype=torch.bfloat16, device=dev)
    calls.mae_gather_rows_bf16(p(src), None, B, L - 1, L, 1, D, p(out2), stream())
    assert torch.equal(out2, bf(src[:, 1:].reshape(-1, D)))
    acc = torch.zeros(D, device=dev)
    calls.mae_colsum_rows_f32(p(src), p(m), B, L - Lk, L, 0, D, p(acc), stream())
    assert max_err(acc, O.get_at_index(src, m).sum((0, 1))) < 1e-3
    acc.zero_()
    calls.mae_colsum_rows_f32(p(src), None, B, 1, L, 0, D, p(acc), stream())
    assert max_err(acc, src[:, 0].sum(0)) < 1e-4
    xb = bf(torch.randn(3001, 432, generator=g)).to(dev)
    cs = torch.zeros(432, device=dev)
    calls.mae_colsum_bf16(p(xb), 3001, 432, p(cs), stream())
    assert rel_err(cs, xb.float().sum(0)) < 1e-5


@pytest.mark.parametrize("norm_pix", [0, 1])
def test_masked_mse(norm_pix):
    B, L, Lk = 9, 145, 36
    g = torch.Generator().manual_seed(21)
    images = torch.randn(B, 3, 96, 96, generator=g).to(dev)
    _, m = O.random_token_mask_from_noise(torch.rand(B, L, generator=g), 0.75)
    m = m.to(dev)
    Lm = L - Lk
    pred = bf(torch.randn(B, Lm, 192, generator=g)).to(dev)
    loss, parts = torch.zeros(1, device=dev), torch.empty(4096, device=dev)
    dpred = torch.empty_like(pred)
    target = torch.empty(B, Lm, 192, device=dev)
    calls.mae_masked_mse(p(pred), p(images), p(m), B, 3, 96, 8, Lm, norm_pix, p(loss), p(parts), p(dpred), p(target), stream())
    t_ref = O.get_at_index(O.patchify(images, 8), torch.clamp(m - 1, min=0))
    if norm_pix:
        t_ref = (t_ref - t_ref.mean(-1, keepdim=True)) / (t_ref.var(-1, keepdim=True) + 1e-6) ** 0.5
        assert max_err(target, t_ref) < 1e-4
    else:
        assert torch.equal(target, t_ref)  # pure data movement: exact
    pr = pred.float().requires_grad_(True)
    l_ref = F.mse_loss(pr, t_ref)
    l_ref.backward()
    assert abs(loss.item() - l_ref.item()) < 1e-5 * l_ref.item() + 1e-7
    assert max_err(dpred, pr.grad) <= 2 ** -8 * pr.grad.abs().max().item()


def test_gemm_tn_gelu_table_exact():
    """EPI 1 on hand-picked pre-activations (identity weights, K = 16): the table-driven GELU / GELU' must equal the
    correctly rounded bf16 of the double-precision erf formulas for EVERY bf16 input, including the out-of-table
    inputs (|x| < 2^-20, |x| > 16, +-0, +-inf)."""
    K = N = 16
    bits = torch.arange(0, 65536, dtype=torch.int32)
    a = bits & 0x7fff
    bits = bits[(a < 0x7f80) & ((a >= 0x80) | (a == 0))]  # finite, no subnormals (inf / nan are fed through the bias below)
    xs = bits.to(torch.int16).view(torch.bfloat16)
    M = (xs.numel() + N - 1) // N
    pad = torch.zeros(M * N - xs.numel(), dtype=torch.bfloat16)
    X = torch.cat([xs, pad]).view(M, N)
    # identity weights: acc[m, n] = X[m, n] exactly, so every bf16 value reaches the epilogue as a pre-activation
    A = X.to(dev)
    I = bf(torch.eye(N)).to(dev)
    h, dact = gemm_tn(A, I, 1, None)

    def refs(xd):
        cdf = 0.5 * torch.erfc(-xd / math.sqrt(2.0))
        pdf = torch.exp(-0.5 * xd * xd) / math.sqrt(2 * math.pi)
        h_ref = torch.where(torch.isinf(xd), torch.where(xd > 0, xd, torch.zeros_like(xd)), xd * cdf)
        d_ref = torch.where(torch.isinf(xd), (xd > 0).double(), cdf + xd * pdf)
        return h_ref.float().to(torch.bfloat16), d_ref.float().to(torch.bfloat16)

    h_ref, d_ref = refs(X.double())
    hc, dc = h.cpu(), dact.cpu()
    assert torch.equal(hc.float(), h_ref.float()), (hc.float() - h_ref.float()).abs().max()
    assert torch.equal(dc.float(), d_ref.float()), (dc.float() - d_ref.float()).abs().max()
    # ... and it agrees with torch's own fp32 erf GELU of the bf16 input to within one bf16 ulp, plus the fp32
    # cancellation error of torch's 1 + erf(x / sqrt 2) in the negative tail (~6e-8 |x|; the table has none)
    t = F.gelu(X.float()).to(torch.bfloat16).float()
    ulp = torch.maximum(t.abs(), torch.tensor(2.0 ** -120)) * 2 ** -7
    assert ((hc.float() - t).abs() <= ulp + 2e-7 * X.float().abs()).all()
    report("gemm_tn_gelu_table", mismatches_vs_torch=int((hc.float() != t).sum()), n=int(t.numel()))
    # +-inf, +-big through the bias (acc = 0)
    special = torch.tensor([float("inf"), float("-inf"), 3e38, -3e38, 17.0, -17.0, 16.0, -16.0, 1e-7, -1e-7, 2.0 ** -20, -2.0 ** -20,
                            0.0, -0.0, 1.0, -1.0])
    hs, ds = gemm_tn(torch.zeros(64, K, dtype=torch.bfloat16, device=dev), I, 1, special.to(dev))
    hs_ref, ds_ref = refs(special.to(torch.bfloat16).double())
    assert torch.equal(hs[0].cpu().float(), hs_ref.float()) and torch.equal(ds[5].cpu().float(), ds_ref.float())
    nan = torch.full((128, 16), float("nan"), dtype=torch.bfloat16, device=dev)
    hn, dn = gemm_tn(nan, I, 1, None)
    assert torch.isnan(hn.float()).all() and torch.isnan(dn.float()).all()


def test_zero_rows():
    B, L, Lk, D = 7, 145, 36, 192
    g = torch.Generator().manual_seed(12)
    k, m = [t.to(dev) for t in O.random_token_mask_from_noise(torch.rand(B, L, generator=g), 0.75)]
    x = torch.randn(B, L, D, generator=g).to(dev)
    xb = bf(x)
    x0, xb0 = x.clone(), xb.clone()
    calls.mae_zero_rows(p(x), p(xb), p(k), B, Lk, L, D, stream())
    keep = torch.zeros(B, L, dtype=torch.bool, device=dev).scatter_(1, k, True)
    assert (x[keep] == 0).all() and (xb[keep] == 0).all()
    assert torch.equal(x[~keep], x0[~keep]) and torch.equal(xb[~keep], xb0[~keep])


@pytest.mark.parametrize("patch,img,norm_pix", [(4, 32, 0), (16, 64, 1), (6, 36, 0)])
def test_masked_mse_generic_patch(patch, img, norm_pix):
    """patch sizes other than the 8x8 fast path go through the generic kernel."""
    B = 5
    n_patches = (img // patch) ** 2
    L, pd = n_patches + 1, 3 * patch * patch
    g = torch.Generator().manual_seed(22)
    images = torch.randn(B, 3, img, img, generator=g).to(dev)
    _, m = O.random_token_mask_from_noise(torch.rand(B, L, generator=g), 0.75)
    m = m.to(dev)
    Lm = m.shape[1]
    pred = bf(torch.randn(B, Lm, pd, generator=g)).to(dev)
    loss, parts = torch.zeros(1, device=dev), torch.empty(4096, device=dev)
    dpred, target = torch.empty_like(pred), torch.empty(B, Lm, pd, device=dev)
    calls.mae_masked_mse(p(pred), p(images), p(m), B, 3, img, patch, Lm, norm_pix, p(loss), p(parts), p(dpred), p(target), stream())
    t_ref = O.get_at_index(O.patchify(images, patch), torch.clamp(m - 1, min=0))
    if norm_pix:
        t_ref = (t_ref - t_ref.mean(-1, keepdim=True)) / (t_ref.var(-1, keepdim=True) + 1e-6) ** 0.5
        assert max_err(target, t_ref) < 1e-4
    else:
        assert torch.equal(target, t_ref)
    l_ref = F.mse_loss(pred.float(), t_ref)
    assert abs(loss.item() - l_ref.item()) < 1e-5 * l_ref.item() + 1e-7


def test_casts_and_clip_adamw():
    g = torch.Generator().manual_seed(31)
    n = 100_003
    src = torch.randn(n, generator=g).to(dev)
    dst = torch.empty(n, dtype=torch.bfloat16, device=dev)
    calls.mae_cast_bf16(p(src), p(dst), n, stream())
    assert torch.equal(dst, bf(src))
    o3 = 4096 + 432 * 144 + 64
    flat = torch.randn(o3 + 3072 * 768, generator=g).to(dev)
    tbl = torch.tensor([[0, 48, 64, 0], [4096, 432, 144, 4096], [o3, 3072, 768, o3]], dtype=torch.int64, device=dev)
    dt = torch.zeros(flat.numel(), dtype=torch.bfloat16, device=dev)
    calls.mae_transpose_cast_bf16(p(flat), p(dt), p(tbl), 3, 3072, 768, stream())
    assert torch.equal(dt[:48 * 64].view(64, 48), bf(flat[:48 * 64].view(48, 64).t()))
    assert torch.equal(dt[4096:4096 + 432 * 144].view(144, 432), bf(flat[4096:4096 + 432 * 144].view(432, 144).t()))
    assert torch.equal(dt[o3:].view(768, 3072), bf(flat[o3:].view(3072, 768).t()))
    # clip + AdamW, three steps, against the oracle's explicit restatement
    pr = torch.randn(n, generator=g)
    gr = torch.randn(n, generator=g) * 0.01
    pd_, m_, v_ = pr.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    shadow = torch.empty(n, dtype=torch.bfloat16, device=dev)
    parts, gn = torch.zeros(1024, device=dev), torch.zeros(1, device=dev)
    rp, rm_, rv = [pr.clone()], [torch.zeros(n)], [torch.zeros(n)]
    for step in (1, 2, 3):
        gstep = gr * step * 40  # crosses the clip threshold
        gd = gstep.to(dev)
        calls.mae_grad_sumsq(p(gd), n, p(parts), 1024, stream())
        calls.mae_clip_adamw(p(pd_), p(gd), p(m_), p(v_), p(shadow), n, 1e-3, 0.9, 0.999, 1e-8, 0.05, 1.0, step, p(parts), 1024, p(gn), stream())
        total = O.clip_adamw_reference(rp, [gstep], rm_, rv, step, lr=1e-3)
        assert abs(gn.item() - total.item()) < 1e-4 * total.item()
        torch.cuda.synchronize()
    assert max_err(pd_.cpu(), rp[0]) < 1e-6 and torch.equal(shadow, bf(pd_))
