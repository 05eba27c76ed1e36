"""``torch.library`` custom ops over the C ABI: ``torch.ops.mae_b200.*`` (BASELINE.json ``north_star``: the Python modules
"call PyTorch custom ops through a thin C++/C-ABI extension"; SURVEY §8b layer 1).

Every op is registered for the CUDA dispatch key only -- a CPU tensor raises ``NotImplementedError`` from the dispatcher:
there is no CPU fallback -- takes / returns ``torch.Tensor`` s, checks dtype / shape / contiguity, fetches the current CUDA
stream and forwards raw pointers + sizes to the ``extern "C"`` entry point of the same name in ``libmae_b200.so``
(``include/mae_b200.h``).  Fake (meta) implementations make them traceable; the differentiable building blocks
(``linear``, ``layernorm``, ``attention``, ``masked_mse``) carry ``register_autograd`` formulas that call the backward
kernels, so a user module written against these ops trains through plain ``loss.backward()``.

The engine (``mae.py``) sequences the same entry points directly on flat buffers and a preallocated workspace (no per-op
allocation, no autograd graph); these ops are the public, composable surface of the same kernels.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._lib import EPI_BF16, EPI_RESID_F32, calls

NS = "mae_b200"
bf16, f32 = torch.bfloat16, torch.float32


def _st(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _chk(t: torch.Tensor, dtype, ndim: Optional[int] = None, name: str = "tensor") -> torch.Tensor:
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected {dtype}, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got {tuple(t.shape)}")
    return t.contiguous()


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# --------------------------------------------------------------------------------------
# raw kernels (one op per C entry point that a module author composes)
# --------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::random_token_mask", mutates_args=(), device_types="cuda")
def random_token_mask(noise: torch.Tensor, num_keep: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """lightly ``random_token_mask`` given the ``torch.rand`` draw: ``(idx_keep (B,num_keep), idx_mask (B,L-num_keep))``."""
    noise = _chk(noise, f32, 2, "noise")
    B, L = noise.shape
    keep = torch.empty(B, num_keep, dtype=torch.int64, device=noise.device)
    mask = torch.empty(B, L - num_keep, dtype=torch.int64, device=noise.device)
    calls.mae_random_token_mask(noise.data_ptr(), B, L, num_keep, keep.data_ptr(), mask.data_ptr(), _st(noise))
    return keep, mask


@random_token_mask.register_fake
def _(noise, num_keep):
    B, L = noise.shape
    return noise.new_empty((B, num_keep), dtype=torch.int64), noise.new_empty((B, L - num_keep), dtype=torch.int64)


@torch.library.custom_op(f"{NS}::gemm_tn", mutates_args=(), device_types="cuda")
def gemm_tn(a: torch.Tensor, b: torch.Tensor, bias: Optional[torch.Tensor], epi: int,
            aux: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """``epi(a[M,K] @ b[N,K]^T + bias)`` on tcgen05 (``mae_gemm_tn``).  Returns ``(out, second)``: ``second`` is GELU'
    for ``epi == 1`` and an empty tensor otherwise."""
    a, b = _chk(a, bf16, 2, "a"), _chk(b, bf16, 2, "b")
    M, K = a.shape
    N = b.shape[0]
    if b.shape[1] != K:
        raise ValueError("gemm_tn: a is (M,K), b must be (N,K)")
    out = torch.empty(M, N, dtype=f32 if epi == EPI_RESID_F32 else bf16, device=a.device)
    second = torch.empty((M, N) if epi == 1 else (0,), dtype=bf16, device=a.device)
    calls.mae_gemm_tn(a.data_ptr(), b.data_ptr(), M, N, K, epi, _p(None if bias is None else _chk(bias, f32, 1, "bias")),
                      out.data_ptr(), second.data_ptr() if epi == 1 else None,
                      _p(None if aux is None else aux.contiguous()), _st(a))
    return out, second


@gemm_tn.register_fake
def _(a, b, bias, epi, aux):
    M, N = a.shape[0], b.shape[0]
    return a.new_empty((M, N), dtype=f32 if epi == EPI_RESID_F32 else bf16), a.new_empty((M, N) if epi == 1 else (0,), dtype=bf16)


@torch.library.custom_op(f"{NS}::gemm_wgrad", mutates_args=(), device_types="cuda")
def gemm_wgrad(dy: torch.Tensor, x: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """``(dW (n_out, n_in) f32 = dy^T x, db (n_out) f32 = colsum(dy))`` in one launch (``mae_gemm_wgrad``)."""
    dy, x = _chk(dy, bf16, 2, "dy"), _chk(x, bf16, 2, "x")
    M, n_out = dy.shape
    n_in = x.shape[1]
    dw = torch.zeros(n_out, n_in, dtype=f32, device=dy.device)
    db = torch.zeros(n_out, dtype=f32, device=dy.device)
    if n_in <= n_out:  # the wider operand rides the MMA M axis
        calls.mae_gemm_wgrad(dy.data_ptr(), n_out, x.data_ptr(), n_in, M, dw.data_ptr(), n_in, 1, db.data_ptr(), 0, _st(dy))
    else:
        calls.mae_gemm_wgrad(x.data_ptr(), n_in, dy.data_ptr(), n_out, M, dw.data_ptr(), 1, n_in, db.data_ptr(), 1, _st(dy))
    return dw, db


@gemm_wgrad.register_fake
def _(dy, x):
    return dy.new_empty((dy.shape[1], x.shape[1]), dtype=f32), dy.new_empty((dy.shape[1],), dtype=f32)


@torch.library.custom_op(f"{NS}::patch_embed_tokens", mutates_args=(), device_types="cuda")
def patch_embed_tokens(images: torch.Tensor, idx_keep: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor,
                       cls_token: torch.Tensor, pos_embed: torch.Tensor, patch: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Kept-token patch embedding in one launch (``mae_patch_embed_tokens``): ``(x (B,num_keep,D) f32, patches bf16)``."""
    images, idx_keep = _chk(images, f32, 4, "images"), _chk(idx_keep, torch.int64, 2, "idx_keep")
    B, C, img, _ = images.shape
    D, nk = weight.shape[0], idx_keep.shape[1]
    w = _chk(weight.reshape(D, -1), bf16, 2, "weight")
    x = torch.empty(B, nk, D, dtype=f32, device=images.device)
    patches = torch.empty(B * (nk - 1), C * patch * patch, dtype=bf16, device=images.device)
    calls.mae_patch_embed_tokens(images.data_ptr(), idx_keep.data_ptr(), B, C, img, patch, nk, D, w.data_ptr(),
                                 _chk(bias, f32, 1, "bias").data_ptr(), _chk(cls_token.reshape(-1), f32, 1).data_ptr(),
                                 _chk(pos_embed.reshape(-1, D), f32, 2).data_ptr(), patches.data_ptr(), x.data_ptr(), _st(images))
    return x, patches


@patch_embed_tokens.register_fake
def _(images, idx_keep, weight, bias, cls_token, pos_embed, patch):
    B, C = images.shape[:2]
    nk, D = idx_keep.shape[1], weight.shape[0]
    return images.new_empty((B, nk, D)), images.new_empty((B * (nk - 1), C * patch * patch), dtype=bf16)


@torch.library.custom_op(f"{NS}::mlp_fwd", mutates_args=(), device_types="cuda")
def mlp_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, w1: torch.Tensor, b1: torch.Tensor, w2: torch.Tensor,
            b2: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``x + fc2(GELU(fc1(LayerNorm(x))))`` in one launch (``mae_mlp_fwd``): ``(y, xn, mean, rstd, h, gelu')``."""
    x = _chk(x, f32, 2, "x")
    M, D = x.shape
    dev = x.device
    y, xn = torch.empty(M, D, device=dev), torch.empty(M, D, dtype=bf16, device=dev)
    mean, rstd = torch.empty(M, device=dev), torch.empty(M, device=dev)
    h, dg = torch.empty(M, 4 * D, dtype=bf16, device=dev), torch.empty(M, 4 * D, dtype=bf16, device=dev)
    calls.mae_mlp_fwd(x.data_ptr(), None, 0, 0, M, D, _chk(gamma, f32).data_ptr(), _chk(beta, f32).data_ptr(), float(eps),
                      _chk(w1, bf16, 2).data_ptr(), _chk(b1, f32).data_ptr(), _chk(w2, bf16, 2).data_ptr(),
                      _chk(b2, f32).data_ptr(), xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(), h.data_ptr(), dg.data_ptr(),
                      y.data_ptr(), _st(x))
    return y, xn, mean, rstd, h, dg


@mlp_fwd.register_fake
def _(x, gamma, beta, w1, b1, w2, b2, eps):
    M, D = x.shape
    return (x.new_empty((M, D)), x.new_empty((M, D), dtype=bf16), x.new_empty((M,)), x.new_empty((M,)),
            x.new_empty((M, 4 * D), dtype=bf16), x.new_empty((M, 4 * D), dtype=bf16))


# --------------------------------------------------------------------------------------
# differentiable building blocks
# --------------------------------------------------------------------------------------
@torch.library.custom_op(f"{NS}::linear", mutates_args=(), device_types="cuda")
def linear(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """``nn.Linear`` under autocast: bf16 ``x (M,K)``, bf16 ``weight (N,K)``, fp32 ``bias`` -> bf16 ``(M,N)``."""
    out, _ = gemm_tn(x, weight, bias, EPI_BF16, None)
    return out


@linear.register_fake
def _(x, weight, bias):
    return x.new_empty((x.shape[0], weight.shape[0]), dtype=bf16)


def _linear_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1])


def _linear_bwd(ctx, dy):
    x, w = ctx.saved_tensors
    dy = dy.to(bf16).contiguous()
    dx, _ = gemm_tn(dy, w.t().contiguous(), None, EPI_BF16, None)   # dgrad: B operand = W^T, K-major
    dw, db = gemm_wgrad(dy, x)                                       # fp32 weight / bias gradients, one launch
    return dx, dw.to(w.dtype), db


linear.register_autograd(_linear_bwd, setup_context=_linear_setup)


@torch.library.custom_op(f"{NS}::layernorm", mutates_args=(), device_types="cuda")
def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """``nn.LayerNorm(D, eps)`` of an fp32 ``(rows, D)`` stream -> ``(xn bf16, mean f32, rstd f32)``."""
    x = _chk(x, f32, 2, "x")
    M, D = x.shape
    xn = torch.empty(M, D, dtype=bf16, device=x.device)
    mean, rstd = torch.empty(M, device=x.device), torch.empty(M, device=x.device)
    calls.mae_layernorm_fwd(x.data_ptr(), _chk(gamma, f32).data_ptr(), _chk(beta, f32).data_ptr(), None, 0, 0, M, D, float(eps),
                            xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _st(x))
    return xn, mean, rstd


@layernorm.register_fake
def _(x, gamma, beta, eps):
    return x.new_empty(x.shape, dtype=bf16), x.new_empty((x.shape[0],)), x.new_empty((x.shape[0],))


@torch.library.custom_op(f"{NS}::layernorm_bwd", mutates_args=(), device_types="cuda")
def layernorm_bwd(g: torch.Tensor, x: torch.Tensor, mean: torch.Tensor, rstd: torch.Tensor,
                  gamma: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    g, x = _chk(g, bf16, 2, "g"), _chk(x, f32, 2, "x")
    M, D = x.shape
    dx = torch.empty(M, D, device=x.device)
    dgamma, dbeta = torch.zeros(D, device=x.device), torch.zeros(D, device=x.device)
    calls.mae_layernorm_bwd(g.data_ptr(), x.data_ptr(), mean.data_ptr(), rstd.data_ptr(), _chk(gamma, f32).data_ptr(), None, None,
                            0, 0, M, D, dx.data_ptr(), None, dgamma.data_ptr(), dbeta.data_ptr(), _st(x))
    return dx, dgamma, dbeta


@layernorm_bwd.register_fake
def _(g, x, mean, rstd, gamma):
    return x.new_empty(x.shape), x.new_empty((x.shape[1],)), x.new_empty((x.shape[1],))


def _ln_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], inputs[1], output[1], output[2])


def _ln_bwd(ctx, g, _gm, _gr):
    x, gamma, mean, rstd = ctx.saved_tensors
    dx, dgamma, dbeta = layernorm_bwd(g.to(bf16).contiguous(), x, mean, rstd, gamma)
    return dx, dgamma, dbeta, None


layernorm.register_autograd(_ln_bwd, setup_context=_ln_setup)


@torch.library.custom_op(f"{NS}::attention", mutates_args=(), device_types="cuda")
def attention(qkv: torch.Tensor, num_heads: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """timm ``Attention`` core on ``qkv (B, L, 3*H*hd)`` bf16 -> ``(out (B, L, H*hd) bf16, lse (B, H, L) f32)``."""
    qkv = _chk(qkv, bf16, 3, "qkv")
    B, L, D3 = qkv.shape
    D = D3 // 3
    out = torch.empty(B, L, D, dtype=bf16, device=qkv.device)
    lse = torch.empty(B, num_heads, L, dtype=f32, device=qkv.device)
    calls.mae_attention_fwd(qkv.data_ptr(), B, L, num_heads, D // num_heads, out.data_ptr(), lse.data_ptr(), _st(qkv))
    return out, lse


@attention.register_fake
def _(qkv, num_heads):
    B, L, D3 = qkv.shape
    return qkv.new_empty((B, L, D3 // 3)), qkv.new_empty((B, num_heads, L), dtype=f32)


@torch.library.custom_op(f"{NS}::attention_bwd", mutates_args=(), device_types="cuda")
def attention_bwd(qkv: torch.Tensor, out: torch.Tensor, dout: torch.Tensor, lse: torch.Tensor, num_heads: int) -> torch.Tensor:
    B, L, D3 = qkv.shape
    dqkv = torch.empty_like(qkv)
    calls.mae_attention_bwd(qkv.data_ptr(), out.data_ptr(), _chk(dout, bf16, 3, "dout").data_ptr(), lse.data_ptr(), B, L, num_heads,
                            D3 // 3 // num_heads, dqkv.data_ptr(), _st(qkv))
    return dqkv


@attention_bwd.register_fake
def _(qkv, out, dout, lse, num_heads):
    return torch.empty_like(qkv)


def _attn_setup(ctx, inputs, output):
    ctx.save_for_backward(inputs[0], output[0], output[1])
    ctx.num_heads = inputs[1]


def _attn_bwd(ctx, dout, _dlse):
    qkv, out, lse = ctx.saved_tensors
    return attention_bwd(qkv, out, dout.to(bf16).contiguous(), lse, ctx.num_heads), None


attention.register_autograd(_attn_bwd, setup_context=_attn_setup)


@torch.library.custom_op(f"{NS}::masked_mse", mutates_args=(), device_types="cuda")
def masked_mse(pred: torch.Tensor, images: torch.Tensor, idx_mask: torch.Tensor, patch: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """``MSELoss(pred, get_at_index(patchify(images), idx_mask - 1))`` -> ``(loss (), dloss/dpred bf16)`` (``mae_masked_mse``)."""
    pred, images, idx_mask = _chk(pred, bf16, 3, "pred"), _chk(images, f32, 4, "images"), _chk(idx_mask, torch.int64, 2)
    B, C, img, _ = images.shape
    loss = torch.zeros(1, dtype=f32, device=pred.device)
    partials = torch.empty(4096, dtype=f32, device=pred.device)
    dpred = torch.empty_like(pred)
    calls.mae_masked_mse(pred.data_ptr(), images.data_ptr(), idx_mask.data_ptr(), B, C, img, patch, idx_mask.shape[1], 0,
                         loss.data_ptr(), partials.data_ptr(), dpred.data_ptr(), None, _st(pred))
    return loss[0].clone(), dpred


@masked_mse.register_fake
def _(pred, images, idx_mask, patch):
    return pred.new_empty((), dtype=f32), torch.empty_like(pred)


def _mse_setup(ctx, inputs, output):
    ctx.save_for_backward(output[1])


def _mse_bwd(ctx, dloss, _d):
    (dpred,) = ctx.saved_tensors
    return (dpred.float() * dloss).to(bf16), None, None, None


masked_mse.register_autograd(_mse_bwd, setup_context=_mse_setup)

__all__ = ["random_token_mask", "gemm_tn", "gemm_wgrad", "patch_embed_tokens", "mlp_fwd", "linear", "layernorm",
           "layernorm_bwd", "attention", "attention_bwd", "masked_mse"]
