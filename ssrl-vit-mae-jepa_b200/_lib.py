"""ctypes binding of ``libmae_b200.so`` (the C ABI declared in ``include/mae_b200.h``).

There is no CPU fallback: if the shared library is missing the loader raises, and every entry point
returns a negative error code (raised here as ``RuntimeError``) instead of silently computing elsewhere.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmae_b200.so")

_vp, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float

# name -> argtypes (stream is always last).  Mirrors include/mae_b200.h one to one.
SIGNATURES = {
    "mae_random_token_mask": [_vp, _i, _i, _i, _vp, _vp, _vp],
    "mae_gather_patches": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "mae_patch_embed_tokens": [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mae_assemble_tokens": [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp],
    "mae_layernorm_fwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp],
    "mae_layernorm_bwd": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "mae_layernorm_fwd_add": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _f, _vp, _vp, _vp, _vp, _vp],
    "mae_layernorm_bwd_scatter": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "mae_gemm_tn": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "mae_mlp_fwd": [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "mae_gemm_wgrad": [_vp, _i, _vp, _i, _i, _vp, _ll, _ll, _vp, _i, _vp],
    "mae_attention_fwd": [_vp, _i, _i, _i, _i, _vp, _vp, _vp],
    "mae_attention_bwd": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "mae_attention_bwd_ws": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _ll, _vp],
    "mae_attention_fwd_q": [_vp, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "mae_attention_bwd_q": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "mae_scatter_decoder_tokens": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "mae_scatter_decoder_tokens_grouped": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp],
    "mae_scatter_decoder_tokens_ln": [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _f, _vp, _vp, _vp, _vp, _vp],
    "mae_masked_mse": [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp],
    "mae_colsum_bf16": [_vp, _i, _i, _vp, _vp],
    "mae_gather_rows_bf16": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "mae_colsum_rows_f32": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "mae_zero_rows": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "mae_augment_u8": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "mae_augment_params": [_vp, _i, _i, _i, _f, _f, _f, _f, C.c_ulonglong, C.c_ulonglong, _i, _vp],
    "mae_cast_bf16": [_vp, _vp, _ll, _vp],
    "mae_transpose_cast_bf16": [_vp, _vp, _vp, _i, _i, _i, _vp],
    "mae_allreduce_avg_sumsq": [_vp, _i, _ll, _ll, _vp, _vp, _i, _vp],
    "mae_grad_sumsq": [_vp, _ll, _vp, _i, _vp],
    "mae_clip_adamw": [_vp, _vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _f, _i, _vp, _i, _vp, _vp],
}

EPI_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_DGELU_BF16 = 0, 1, 2, 4

_lib: Optional[C.CDLL] = None


def load(build_if_missing: bool = True) -> C.CDLL:
    """Load (building first if needed) the shared library and attach argtypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{LIB_PATH} is missing and there is no CPU fallback; run build.py")
        from . import build as _build
        _build.build()
    lib = C.CDLL(LIB_PATH)
    lib.mae_version.restype = _i
    lib.mae_version.argtypes = []
    lib.mae_error_string.restype = C.c_char_p
    lib.mae_error_string.argtypes = [_i]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = _i
        fn.argtypes = argtypes
    _lib = lib
    return lib


class _Calls:
    """``calls.mae_xxx(...)``: invoke the C entry point, count the launch, raise on a non-zero code."""

    def __init__(self):
        self.launches = 0

    def __getattr__(self, name: str):
        lib = load()
        fn = getattr(lib, name)

        def call(*args):
            rc = fn(*args)
            if rc != 0:
                raise RuntimeError(f"{name} failed: {lib.mae_error_string(rc).decode()} (code {rc})")
            self.launches += 1

        setattr(self, name, call)
        return call


calls = _Calls()
