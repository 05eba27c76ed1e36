"""Host-side mirror of the reference's ``src/models/mae.py`` on top of ``libmae_b200.so``.

``MaskedAutoencoder`` keeps the reference's constructor (three config dicts with the same ``.get``
defaults, ``src/models/mae.py:15-52``), attributes, ``forward(images) -> (x_pred, target)`` contract
(``:77-94``), ``forward_encoder`` / ``forward_decoder`` (``:54-75``) and -- parameter for parameter --
the 87-tensor ``state_dict`` layout of the lightly/timm module tree it replaces, so ``vit-mae.pt``
checkpoints load either way.  The sub-modules are parameter holders only: every FLOP runs in the
hand-written sm_100a kernels through :class:`MAEEngine`.  CPU tensors raise; there is no fallback.
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _lib
from ._lib import EPI_BF16, EPI_DGELU_BF16, EPI_GELU_BF16, EPI_RESID_F32, calls

LN_EPS = 1e-6  # timm VisionTransformer / lightly MAEDecoderTIMM use LayerNorm(eps=1e-6)


# --------------------------------------------------------------------------------------
# parameter-holder module tree (names == timm 1.0.21 / lightly 1.5.22 state_dict keys)
# --------------------------------------------------------------------------------------
class _HolderMixin:
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("this sub-module only holds parameters; run it through MaskedAutoencoder / MAEEngine")


class _Attention(_HolderMixin, nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads, self.head_dim = num_heads, dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(_HolderMixin, nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(_HolderMixin, nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = _Attention(dim, num_heads)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))


class _PatchEmbed(_HolderMixin, nn.Module):
    def __init__(self, img_size: int, patch_size: int, in_chans: int, embed_dim: int):
        super().__init__()
        self.img_size, self.patch_size, self.in_chans = img_size, patch_size, in_chans
        self.grid_size = img_size // patch_size
        self.num_patches = self.grid_size**2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)


def sincos_2d(embed_dim: int, grid_size: int) -> torch.Tensor:
    """Frozen 2-D sin-cos position table with a zero class-token row: the first half of the
    channels encodes the column, the second half the row."""
    assert embed_dim % 4 == 0
    q = embed_dim // 4
    omega = 1.0 / (10000.0 ** (torch.arange(q, dtype=torch.float64) / q))
    pos = torch.arange(grid_size, dtype=torch.float64)
    ang = pos[:, None] * omega[None, :]                      # (grid, q)
    one = torch.cat([ang.sin(), ang.cos()], dim=1)           # (grid, D/2)
    col = one[None, :, :].expand(grid_size, grid_size, -1)   # varies with the column index
    row = one[:, None, :].expand(grid_size, grid_size, -1)   # varies with the row index
    table = torch.cat([col, row], dim=2).reshape(grid_size * grid_size, embed_dim)
    return torch.cat([torch.zeros(1, embed_dim, dtype=torch.float64), table], dim=0).float().unsqueeze(0)


def _init_linear_ln(root: nn.Module) -> None:
    for m in root.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


class _VisionTransformer(_HolderMixin, nn.Module):
    """Holder with timm ``VisionTransformer`` parameter names (``src/models/mae.py:28-36``).  Also usable on its
    own like ``timm.models.vision_transformer.VisionTransformer(img_size, patch_size, in_chans, embed_dim, depth,
    num_heads, num_classes=0)`` (reference ``src/training/classifier.py:47-55``): ``forward_features`` runs the
    B200 kernels over all tokens (see ``classifier.py``)."""

    def __init__(self, img_size=96, patch_size=8, in_chans=3, embed_dim=144, depth=4, num_heads=6, num_classes=0):
        super().__init__()
        if num_classes != 0:
            raise NotImplementedError("only the headless ViT (num_classes=0) used by the reference is provided")
        self.embed_dim, self.num_heads, self.num_prefix_tokens = embed_dim, num_heads, 1
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, embed_dim) * 0.02)
        self.blocks = nn.Sequential(*[_Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=LN_EPS)
        for m in self.modules():  # timm defaults for a stand-alone ViT; MaskedAutoencoder re-initialises (lightly)
            if isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                nn.init.zeros_(m.bias)
        nn.init.normal_(self.cls_token, std=1e-6)

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        """patch-embed -> cat CLS -> + pos_embed -> blocks -> norm over ALL tokens: (B, num_patches+1, D) fp32
        (reference ``src/models/classifier.py:48``); differentiable w.r.t. the parameters."""
        from .classifier import vit_forward_features

        return vit_forward_features(self, x)


class _MaskedViT(_HolderMixin, nn.Module):
    """Holder with lightly ``MaskedVisionTransformerTIMM`` names: ``mask_token`` + ``vit.*``."""

    def __init__(self, vit: _VisionTransformer):
        super().__init__()
        self.vit = vit
        self.mask_token = nn.Parameter(torch.zeros(1, 1, vit.embed_dim))  # never used by the MAE path
        w = vit.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.normal_(vit.cls_token, std=0.02)
        _init_linear_ln(vit)
        vit.pos_embed.data.copy_(sincos_2d(vit.embed_dim, vit.patch_embed.grid_size))
        vit.pos_embed.requires_grad = False

    @property
    def sequence_length(self) -> int:
        return self.vit.patch_embed.num_patches + self.vit.num_prefix_tokens


class _MAEDecoder(_HolderMixin, nn.Module):
    """Holder with lightly ``MAEDecoderTIMM`` names (``src/models/mae.py:45-52``)."""

    def __init__(self, num_patches, patch_size, in_chans, embed_dim, decoder_embed_dim, decoder_depth, decoder_num_heads):
        super().__init__()
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.Sequential(*[_Block(decoder_embed_dim, decoder_num_heads) for _ in range(decoder_depth)])
        self.decoder_norm = nn.LayerNorm(decoder_embed_dim, eps=LN_EPS)
        self.decoder_pred = nn.Linear(decoder_embed_dim, patch_size**2 * in_chans, bias=True)
        nn.init.normal_(self.mask_token, std=0.02)
        self.decoder_pos_embed.data.copy_(sincos_2d(decoder_embed_dim, int(round(math.sqrt(num_patches)))))
        _init_linear_ln(self)


def num_keep_tokens(seq_len: int, mask_ratio: float) -> int:
    """lightly ``random_token_mask``: ``max(1, int(L * (1 - mask_ratio)))`` (class token included)."""
    return max(1, int(seq_len * (1 - mask_ratio)))


# --------------------------------------------------------------------------------------
# engine: flat parameter storage + kernel sequencing
# --------------------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


class _BlockW:
    """Byte-offset bookkeeping of one transformer block inside the flat buffers."""
    __slots__ = ("dim", "heads", "hd", "hidden", "n1w", "n1b", "qkvw", "qkvb", "projw", "projb", "n2w", "n2b",
                 "fc1w", "fc1b", "fc2w", "fc2b")


class MAEEngine:
    """Owns the flat fp32 parameter / gradient / AdamW-state buffers (every ``nn.Parameter`` of the model
    is a view into them), their bf16 shadows (W and W^T) and the activation workspace, and sequences the
    C-ABI kernels for forward, backward and the fused clip+AdamW step."""

    ALIGN = 64  # elements; keeps every tensor 128-byte aligned in the bf16 shadow (TMA needs 16)
    NOGRAD = ("encoder.mask_token", "encoder.vit.pos_embed", "decoder.decoder_pos_embed")  # never receive a gradient
    EP = "encoder.vit."  # state_dict prefix of the timm-style ViT inside the model

    def __init__(self, model: nn.Module):
        self.model = model
        self.flat: Optional[torch.Tensor] = None
        self._ws_key = None
        self._shadow_fresh, self._shadow_version = False, None
        self.step_count = 0
        self.norm_pix_loss = False
        # weight-gradient GEMMs run on a second stream: they only feed the optimizer, so they trail the dgrad chain and fill
        # its fill / drain bubbles (measured 8.50 -> 8.23 ms/step); MAE_WGRAD_STREAM=0 keeps everything on one stream
        self.wgrad_side_stream = os.environ.get("MAE_WGRAD_STREAM", "1") != "0"
        # masked-rows-only tail of the last decoder block: decoder_norm / decoder_pred only read the idx_mask rows, so the
        # block's norm2 -> mlp -> residual (and their backward) run on those rows alone (MAE_MASKED_TAIL=0: all rows)
        self.masked_tail = os.environ.get("MAE_MASKED_TAIL", "1") != "0"
        # grouped decoder stream (with the masked tail): rows stored as [mask tokens | kept tokens] per sample, so the rows the
        # prediction head reads come first and the last block's attention only needs its first Lm query rows: the decoder-
        # sized tcgen05 kernels then skip their second query tile (mae_attention_fwd_q / bwd_q).  MAE_GROUPED_DECODER=0:
        # token order
        self.grouped_decoder = os.environ.get("MAE_GROUPED_DECODER", "1") != "0"
        # the small optimizer-only launches (W^T refresh, mask-token column sum, kept-row clearing) also on the side stream
        self.side_extras = os.environ.get("MAE_SIDE_EXTRAS", "1") != "0"
        # decoder-input assembly fused with the first decoder block's norm1: correct, but measured slower than the two
        # launches it replaces (8.27 vs 8.16 ms/step: 80 registers, dependent index -> pos loads), so opt-in only
        self.scatter_ln = os.environ.get("MAE_SCATTER_LN", "0") == "1"
        # norm2 + fc1 + GELU + fc2 + residual as one tcgen05 kernel for D <= 192 (csrc/mlp_fused.cu): 35 % fewer HBM bytes,
        # but its epilogues are instruction-issue / latency bound (profiles/r02_mlp_fused.md).  Measured at B = 2000: faster
        # than the three launches when the stream is L2-resident (encoder, 72 k x 144: 103 vs 117 us), slower on the
        # HBM-sized decoder streams (554 vs 466 us).  Default "auto": fused where rows * D * 4 B fits in ~half of L2.
        # MAE_FUSED_MLP=1 / 0 forces it on / off everywhere.
        self.fused_mlp = {"1": "on", "0": "off"}.get(os.environ.get("MAE_FUSED_MLP", "auto"), "auto")
        # patch gather + patch-embed GEMM + bias / pos / class-token assembly in one launch (MAE_FUSED_PATCH_EMBED=0: three)
        self.fused_patch_embed = os.environ.get("MAE_FUSED_PATCH_EMBED", "1") != "0"
        self._side_skip = False
        self._side: Optional[torch.cuda.Stream] = None
        self._wg_pending: Dict[int, torch.cuda.Event] = {}
        self._wg_events: List[torch.cuda.Event] = []
        self._wg_next_event = 0

    # ---------------- flat storage ----------------
    def _ordered_params(self) -> List[Tuple[str, nn.Parameter]]:
        named = list(self.model.named_parameters())
        nograd = set(self.NOGRAD)
        return [(n, p) for n, p in named if n not in nograd] + [(n, p) for n, p in named if n in nograd]

    def _vit(self):
        return self.model.encoder.vit

    def _flat_is_current(self, device: torch.device) -> bool:
        """True iff every parameter still is the view into ``self.flat`` this engine created (``.to()``, another
        engine flattening a shared sub-module, or a replaced ``.data`` all break that)."""
        if self.flat is None or self.flat.device != device:
            return False
        base = self.flat.data_ptr()
        for n, p in self._param_list:
            if p.data_ptr() != base + 4 * self.offsets[n]:
                return False
        return True

    def _compute_layout(self) -> None:
        """Element offsets of every parameter inside the flat buffers (host-only bookkeeping: no device memory)."""
        params = self._ordered_params()
        self._param_list = params
        off, self.offsets, self.n_opt = 0, {}, None
        for n, p in params:
            if n in self.NOGRAD and self.n_opt is None:
                self.n_opt = off
            self.offsets[n] = off
            off += (p.numel() + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.n_total = off
        if self.n_opt is None:
            self.n_opt = off
        self._index_blocks()

    def _ensure_flat(self, device: torch.device) -> None:
        if self._flat_is_current(device):
            return
        if device.type != "cuda":
            raise RuntimeError("MaskedAutoencoder runs only on CUDA (sm_100a); there is no CPU fallback")
        _lib.load()
        self._compute_layout()
        params, off = self._param_list, self.n_total
        flat = torch.zeros(off, dtype=torch.float32, device=device)
        old_m = getattr(self, "exp_avg", None)
        for n, p in params:
            o = self.offsets[n]
            flat[o:o + p.numel()].copy_(p.data.reshape(-1).to(device=device, dtype=torch.float32))
            p.data = flat[o:o + p.numel()].view(p.shape)
        self.flat = flat
        self.grads = torch.zeros(self.n_opt, dtype=torch.float32, device=device)
        if old_m is None or old_m.numel() != self.n_opt:
            self.exp_avg = torch.zeros(self.n_opt, dtype=torch.float32, device=device)
            self.exp_avg_sq = torch.zeros(self.n_opt, dtype=torch.float32, device=device)
        else:
            self.exp_avg, self.exp_avg_sq = self.exp_avg.to(device), self.exp_avg_sq.to(device)
        self.flat_bf16 = torch.empty(off, dtype=torch.bfloat16, device=device)
        self.flat_bf16_t = torch.zeros(off, dtype=torch.bfloat16, device=device)
        self.partials = torch.zeros(8192, dtype=torch.float32, device=device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=device)
        # transposed shadow table: every Linear weight that needs a dgrad
        rows = []
        for n, p in params:
            if n.endswith(".weight") and p.dim() == 2:
                rows.append([self.offsets[n], p.shape[0], p.shape[1], self.offsets[n]])
        self.t_table = torch.tensor(rows, dtype=torch.int64, device=device)
        self.t_max_rows, self.t_max_cols = max(r[1] for r in rows), max(r[2] for r in rows)
        self._shadow_fresh, self._shadow_version = False, None

    def _o(self, name: str) -> int:
        return self.offsets[name]

    def _index_blocks(self) -> None:
        def blk(prefix: str, dim: int, heads: int) -> _BlockW:
            b = _BlockW()
            b.dim, b.heads, b.hd, b.hidden = dim, heads, dim // heads, dim * 4
            for a, k in [("n1w", "norm1.weight"), ("n1b", "norm1.bias"), ("qkvw", "attn.qkv.weight"),
                         ("qkvb", "attn.qkv.bias"), ("projw", "attn.proj.weight"), ("projb", "attn.proj.bias"),
                         ("n2w", "norm2.weight"), ("n2b", "norm2.bias"), ("fc1w", "mlp.fc1.weight"),
                         ("fc1b", "mlp.fc1.bias"), ("fc2w", "mlp.fc2.weight"), ("fc2b", "mlp.fc2.bias")]:
                setattr(b, a, self._o(prefix + k))
            return b

        m = self.model
        vit = self._vit()
        self.enc_blocks = [blk(f"{self.EP}blocks.{i}.", vit.embed_dim, vit.num_heads) for i in range(len(vit.blocks))]
        if not hasattr(m, "decoder"):
            self.dec_blocks = []
            return
        dd = m.decoder.decoder_embed.out_features
        dh = m.decoder.decoder_blocks[0].attn.num_heads if len(m.decoder.decoder_blocks) else 1
        self.dec_blocks = [blk(f"decoder.decoder_blocks.{i}.", dd, dh) for i in range(len(m.decoder.decoder_blocks))]

    def _param_versions(self):
        """Version counters of the flat buffer and of every parameter view.  Each ``nn.Parameter`` carries its own counter
        (``p.data = flat[...]`` detaches it from ``flat``'s), so ``p.copy_`` / ``load_state_dict`` / ``torch.optim`` steps
        bump the parameter's counter and leave ``flat._version`` untouched."""
        return (self.flat._version, tuple(p._version for _, p in self._param_list))

    def _refresh_shadow(self, stream: int) -> None:
        """bf16 copies of W (same layout) and W^T.  Skipped only for the first forward after the fused optimizer, which
        refreshed both itself, and only if no parameter was written through torch since (version counters).  Every
        other forward rebuilds them (two small launches, ~20 us): writes that bypass the counters (``p.data.add_``,
        kernels on raw pointers) are then always seen, e.g. a torch AdamW fine-tune or ``load_state_dict``."""
        if self._shadow_fresh and self._shadow_version == self._param_versions():
            self._shadow_fresh = False
            return
        self._wg_wait(self.flat_bf16_t.data_ptr())  # a W^T refresh still running on the side stream goes first
        calls.mae_cast_bf16(self.flat.data_ptr(), self.flat_bf16.data_ptr(), self.n_total, stream)
        calls.mae_transpose_cast_bf16(self.flat.data_ptr(), self.flat_bf16_t.data_ptr(), self.t_table.data_ptr(),
                                      self.t_table.shape[0], self.t_max_rows, self.t_max_cols, stream)
        self._shadow_fresh = False

    # pointers into the flat buffers (element offsets -> byte addresses)
    def pf(self, off: int) -> int:
        return self.flat.data_ptr() + 4 * off

    def pg(self, off: int) -> int:
        return self.grads.data_ptr() + 4 * off

    def pw(self, off: int) -> int:
        return self.flat_bf16.data_ptr() + 2 * off

    def pwt(self, off: int) -> int:
        return self.flat_bf16_t.data_ptr() + 2 * off

    # ---------------- workspace ----------------
    def _ensure_ws(self, B: int, Lk: int, device: torch.device) -> None:
        key = (B, Lk, device)
        if self._ws_key == key:
            return
        m = self.model
        L, Lm = m.sequence_length, m.sequence_length - Lk
        De, Dd = m.encoder.vit.embed_dim, m.decoder.decoder_embed.out_features
        pd = m.in_chans * m.patch_size**2
        Me, Md, Mm = B * Lk, B * L, B * Lm
        bf, f32 = torch.bfloat16, torch.float32

        def e(shape, dt):
            return torch.empty(shape, dtype=dt, device=device)

        ws: Dict[str, Any] = {}
        ws["idx_keep"] = e((B, Lk), torch.int64)
        ws["idx_mask"] = e((B, Lm), torch.int64)
        # decoder-stream ROW indices of the masked / kept tokens in the grouped layout (the same for every sample)
        ws["pos_mask"] = torch.arange(Lm, dtype=torch.int64, device=device).repeat(B, 1).contiguous()
        ws["pos_keep"] = (Lm + torch.arange(Lk, dtype=torch.int64, device=device)).repeat(B, 1).contiguous()
        ws["patches"] = e((B * (Lk - 1), pd), bf)
        ws["tok"] = e((B * (Lk - 1), De), bf)

        def block_ws(M, D, H):
            return {"x_in": e((M, D), f32), "xn1": e((M, D), bf), "mean1": e((M,), f32), "rstd1": e((M,), f32),
                    "qkv": e((M, 3 * D), bf), "att": e((M, D), bf), "lse": e((M * H,), f32),
                    "x_mid": e((M, D), f32), "xn2": e((M, D), bf), "mean2": e((M,), f32), "rstd2": e((M,), f32),
                    "pre": e((M, 4 * D), bf), "h": e((M, 4 * D), bf)}

        ws["enc"] = [block_ws(Me, De, b.heads) for b in self.enc_blocks]
        ws["x_enc"] = e((Me, De), f32)
        ws["xnf"], ws["meanf"], ws["rstdf"] = e((Me, De), bf), e((Me,), f32), e((Me,), f32)
        ws["emb"] = e((Me, Dd), bf)
        ws["dec"] = [block_ws(Md, Dd, b.heads) for b in self.dec_blocks]
        ws["x_dec"] = e((Md, Dd), f32)
        ws["xnd"], ws["meand"], ws["rstdd"] = e((Mm, Dd), bf), e((Mm,), f32), e((Mm,), f32)
        ws["dlt"] = e((Mm, Dd), bf)  # last block's mlp.fc2 output on the masked rows (masked_tail)
        ws["pred"] = e((B, Lm, pd), bf)
        ws["dpred"] = e((B, Lm, pd), bf)
        ws["loss"] = torch.zeros(1, dtype=f32, device=device)
        ws["loss_partials"] = e((4096,), f32)
        # backward scratch, shared by all blocks (sized for the decoder, the larger of the two)
        Mx, Dx = max(Me, Md), max(De, Dd)
        ws["dx_a"], ws["dx_b"] = e((Mx * Dx,), f32), e((Mx * Dx,), f32)
        ws["dxb_a"], ws["dxb_b"] = e((Mx * Dx,), bf), e((Mx * Dx,), bf)
        ws["dh"] = e((Mx * Dx * 4,), bf)
        ws["g"] = e((Mx * Dx,), bf)
        ws["datt"] = e((Mx * Dx,), bf)
        ws["dqkv"] = e((Mx * Dx * 3,), bf)
        ws["demb"] = e((Me, Dd), bf)
        ws["dtok"] = e((B * (Lk - 1), De), bf)
        self.ws, self._ws_key = ws, key
        self.dims = dict(B=B, L=L, Lk=Lk, Lm=Lm, De=De, Dd=Dd, pd=pd, Me=Me, Md=Md, Mm=Mm)

    # ---------------- forward pieces ----------------
    def _use_grouped(self) -> bool:
        return self.grouped_decoder and self._use_masked_tail() and not self.scatter_ln

    def _rows_mask(self) -> int:
        """Decoder-stream row indices (B, Lm) of the masked tokens."""
        return _ptr(self.ws["pos_mask"] if self._use_grouped() else self.ws["idx_mask"])

    def _rows_keep(self) -> int:
        return _ptr(self.ws["pos_keep"] if self._use_grouped() else self.ws["idx_keep"])

    def _use_masked_tail(self) -> bool:
        return self.masked_tail and len(self.dec_blocks) > 0 and self.dims["Dd"] <= 256

    def _use_fused_mlp(self, D: int, rows: int = 0) -> bool:
        if self.fused_mlp == "off" or not (64 <= D <= 192 and D % 16 == 0):
            return False
        return self.fused_mlp == "on" or rows * D * 4 <= 64 * 1024 * 1024

    def _block_fwd(self, w: _BlockW, s: Dict[str, torch.Tensor], x_out: Optional[torch.Tensor], M: int, B: int, L: int,
                   st: int, tail: bool = False, skip_norm1: bool = False):
        """One pre-norm transformer block.  ``tail``: the last decoder block; norm2 and the MLP run on the idx_mask rows
        only (compact, Mm rows) and mlp.fc2 leaves its bf16 output in ``ws["dlt"]`` without the residual add, which
        :meth:`_decode` folds into decoder_norm (``mae_layernorm_fwd_add``)."""
        D, H, hd = w.dim, w.heads, w.hd
        if not skip_norm1:  # (the decoder's first block gets xn1 / mean1 / rstd1 from mae_scatter_decoder_tokens_ln)
            calls.mae_layernorm_fwd(_ptr(s["x_in"]), self.pf(w.n1w), self.pf(w.n1b), None, 0, 0, M, D, LN_EPS,
                                    _ptr(s["xn1"]), _ptr(s["mean1"]), _ptr(s["rstd1"]), st)
        calls.mae_gemm_tn(_ptr(s["xn1"]), self.pw(w.qkvw), M, 3 * D, D, EPI_BF16, self.pf(w.qkvb), _ptr(s["qkv"]),
                          None, None, st)
        if tail and self._use_grouped():  # only the masked tokens' rows (first Lm of every sample) are read downstream
            calls.mae_attention_fwd_q(_ptr(s["qkv"]), B, L, H, hd, self.dims["Lm"], _ptr(s["att"]), _ptr(s["lse"]), st)
        else:
            calls.mae_attention_fwd(_ptr(s["qkv"]), B, L, H, hd, _ptr(s["att"]), _ptr(s["lse"]), st)
        calls.mae_gemm_tn(_ptr(s["att"]), self.pw(w.projw), M, D, D, EPI_RESID_F32, self.pf(w.projb),
                          _ptr(s["x_mid"]), None, _ptr(s["x_in"]), st)
        if self._use_fused_mlp(D, self.dims["Mm"] if tail else M):
            # norm2 -> fc1 -> GELU -> fc2 -> + residual in one launch (csrc/mlp_fused.cu); the tail variant gathers the
            # idx_mask rows on the way in and leaves the compact fp32 sums  x_mid[idx_mask] + mlp(...)  in x_dec
            ws, d = self.ws, self.dims
            if tail:
                calls.mae_mlp_fwd(_ptr(s["x_mid"]), self._rows_mask(), d["Lm"], L, d["Mm"], D, self.pf(w.n2w),
                                  self.pf(w.n2b), LN_EPS, self.pw(w.fc1w), self.pf(w.fc1b), self.pw(w.fc2w), self.pf(w.fc2b),
                                  _ptr(s["xn2"]), _ptr(s["mean2"]), _ptr(s["rstd2"]), _ptr(s["h"]), _ptr(s["pre"]),
                                  _ptr(ws["x_dec"]), st)
            else:
                calls.mae_mlp_fwd(_ptr(s["x_mid"]), None, 0, 0, M, D, self.pf(w.n2w), self.pf(w.n2b), LN_EPS,
                                  self.pw(w.fc1w), self.pf(w.fc1b), self.pw(w.fc2w), self.pf(w.fc2b), _ptr(s["xn2"]),
                                  _ptr(s["mean2"]), _ptr(s["rstd2"]), _ptr(s["h"]), _ptr(s["pre"]), _ptr(x_out), st)
            return
        if tail:
            ws, d = self.ws, self.dims
            Mt = d["Mm"]
            calls.mae_layernorm_fwd(_ptr(s["x_mid"]), self.pf(w.n2w), self.pf(w.n2b), self._rows_mask(), d["Lm"], L, Mt,
                                    D, LN_EPS, _ptr(s["xn2"]), _ptr(s["mean2"]), _ptr(s["rstd2"]), st)
            calls.mae_gemm_tn(_ptr(s["xn2"]), self.pw(w.fc1w), Mt, 4 * D, D, EPI_GELU_BF16, self.pf(w.fc1b), _ptr(s["h"]),
                              _ptr(s["pre"]), None, st)
            calls.mae_gemm_tn(_ptr(s["h"]), self.pw(w.fc2w), Mt, D, 4 * D, EPI_BF16, self.pf(w.fc2b), _ptr(ws["dlt"]),
                              None, None, st)
            return
        calls.mae_layernorm_fwd(_ptr(s["x_mid"]), self.pf(w.n2w), self.pf(w.n2b), None, 0, 0, M, D, LN_EPS,
                                _ptr(s["xn2"]), _ptr(s["mean2"]), _ptr(s["rstd2"]), st)
        calls.mae_gemm_tn(_ptr(s["xn2"]), self.pw(w.fc1w), M, 4 * D, D, EPI_GELU_BF16, self.pf(w.fc1b), _ptr(s["h"]),
                          _ptr(s["pre"]), None, st)
        calls.mae_gemm_tn(_ptr(s["h"]), self.pw(w.fc2w), M, D, 4 * D, EPI_RESID_F32, self.pf(w.fc2b), _ptr(x_out),
                          None, _ptr(s["x_mid"]), st)

    def forward(self, images: torch.Tensor, noise: Optional[torch.Tensor], mask_ratio: float,
                want_target: bool = False, compute_loss: bool = True):
        """mask -> encode kept tokens -> decode -> predict masked patches (-> fused MSE + dpred)."""
        m = self.model
        if not images.is_cuda:
            raise RuntimeError("MaskedAutoencoder runs only on CUDA tensors (sm_100a kernels; no CPU fallback)")
        if images.dtype != torch.float32 or images.dim() != 4 or images.shape[1] != m.in_chans or \
                images.shape[2] != m.image_size or images.shape[3] != m.image_size:
            raise ValueError(f"expected float32 images (B,{m.in_chans},{m.image_size},{m.image_size}), got "
                             f"{tuple(images.shape)} {images.dtype}")
        images = images.contiguous()
        dev = images.device
        B, L = images.shape[0], m.sequence_length
        Lk = num_keep_tokens(L, mask_ratio)
        if Lk < 2 or Lk >= L:
            raise ValueError(f"mask_ratio {mask_ratio} keeps {Lk} of {L} tokens; need 2 <= keep < L")
        self._ensure_flat(dev)
        self._ensure_ws(B, Lk, dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._refresh_shadow(st)
        ws, d = self.ws, self.dims
        if noise is None:
            noise = torch.rand(B, L, device=dev)  # same RNG call as lightly.random_token_mask
        noise = noise.to(device=dev, dtype=torch.float32).contiguous()
        if noise.shape != (B, L):
            raise ValueError(f"noise must be ({B},{L})")
        calls.mae_random_token_mask(noise.data_ptr(), B, L, Lk, _ptr(ws["idx_keep"]), _ptr(ws["idx_mask"]), st)
        self._encode(images, ws["idx_keep"], st)
        self._decode(st)
        target = None
        if compute_loss or want_target:
            if want_target:
                target = torch.empty((B, d["Lm"], d["pd"]), dtype=torch.float32, device=dev)
            calls.mae_masked_mse(_ptr(ws["pred"]), images.data_ptr(), _ptr(ws["idx_mask"]), B, m.in_chans, m.image_size,
                                 m.patch_size, d["Lm"], int(self.norm_pix_loss), _ptr(ws["loss"]),
                                 _ptr(ws["loss_partials"]), _ptr(ws["dpred"]), _ptr(target), st)
            calls.launches += 1
        self._images = images
        return ws["pred"], target, ws["loss"]

    def _encode(self, images: torch.Tensor, idx_keep: torch.Tensor, st: int) -> None:
        ws, d = self.ws, self.dims
        pe = self._vit().patch_embed
        B, Lk, De = d["B"], d["Lk"], d["De"]
        o, ep = self._o, self.EP
        x0 = ws["enc"][0]["x_in"] if self.enc_blocks else ws["x_enc"]
        if self.fused_patch_embed and d["pd"] <= 256 and De <= 256 and pe.patch_size % 4 == 0 and \
                (d["pd"] // 64 + (d["pd"] % 64 > 0)) * (De * 128 + 2 * 16384) <= 220 * 1024:
            # gather -> tcgen05 GEMM -> + bias, + pos, class-token rows: one launch (csrc/patch_embed.cu)
            calls.mae_patch_embed_tokens(images.data_ptr(), idx_keep.data_ptr(), B, pe.in_chans, pe.img_size, pe.patch_size,
                                         Lk, De, self.pw(o(ep + "patch_embed.proj.weight")),
                                         self.pf(o(ep + "patch_embed.proj.bias")), self.pf(o(ep + "cls_token")),
                                         self.pf(o(ep + "pos_embed")), _ptr(ws["patches"]), _ptr(x0), st)
        else:
            calls.mae_gather_patches(images.data_ptr(), idx_keep.data_ptr(), B, pe.in_chans, pe.img_size, pe.patch_size, Lk,
                                     _ptr(ws["patches"]), st)
            calls.mae_gemm_tn(_ptr(ws["patches"]), self.pw(o(ep + "patch_embed.proj.weight")), B * (Lk - 1), De,
                              d["pd"], EPI_BF16, self.pf(o(ep + "patch_embed.proj.bias")), _ptr(ws["tok"]), None,
                              None, st)
            calls.mae_assemble_tokens(_ptr(ws["tok"]), self.pf(o(ep + "cls_token")), self.pf(o(ep + "pos_embed")),
                                      idx_keep.data_ptr(), B, Lk, De, _ptr(x0), st)
        for i, w in enumerate(self.enc_blocks):
            x_out = ws["enc"][i + 1]["x_in"] if i + 1 < len(self.enc_blocks) else ws["x_enc"]
            self._block_fwd(w, ws["enc"][i], x_out, d["Me"], B, Lk, st)
        calls.mae_layernorm_fwd(_ptr(ws["x_enc"]), self.pf(o(ep + "norm.weight")), self.pf(o(ep + "norm.bias")),
                                None, 0, 0, d["Me"], De, LN_EPS, _ptr(ws["xnf"]), _ptr(ws["meanf"]), _ptr(ws["rstdf"]), st)

    def grad_bucket_boundaries(self) -> List[int]:
        """Offsets into ``self.grads`` that cut it into one bucket per transformer block, in parameter order: encoder
        bucket 0 = stem + block 0, the last encoder bucket also holds ``vit.norm``; decoder bucket 0 = mask token +
        ``decoder_embed`` + block 0, the last one also holds ``decoder_norm`` / ``decoder_pred``.  Backward completes them
        from the last to the first (see :meth:`backward`)."""
        b = [0] + [w.n1w for w in self.enc_blocks[1:]]
        if self.dec_blocks:
            b += [self._o("decoder.mask_token")] + [w.n1w for w in self.dec_blocks[1:]]
        return b + [self.n_opt]

    def _encoder_bwd(self, g_ptr: int, cur, curb, nxt, nxtb, st: int, first_block: int = 0, stem: bool = True,
                     bucket_cb=None):
        """Backward of ``vit.norm`` + encoder blocks [first_block:] (+ cls_token / patch-embed when ``stem``), given
        the bf16 gradient ``g`` w.r.t. the normalised encoder output.  The result lands in ``cur``-style buffers."""
        ws, d = self.ws, self.dims
        B, Lk, De, pd, Me = d["B"], d["Lk"], d["De"], d["pd"], d["Me"]
        o, ep = self._o, self.EP
        self._wg_wait(self.flat_bf16_t.data_ptr())
        self._wg_wait(_ptr(nxtb))
        self._wg_wait(_ptr(nxt))
        calls.mae_layernorm_bwd(g_ptr, _ptr(ws["x_enc"]), _ptr(ws["meanf"]), _ptr(ws["rstdf"]),
                                self.pf(o(ep + "norm.weight")), None, None, 0, 0, Me, De, _ptr(nxt), _ptr(nxtb),
                                self.pg(o(ep + "norm.weight")), self.pg(o(ep + "norm.bias")), st)
        cur, curb, nxt, nxtb = nxt, nxtb, cur, curb
        for i in reversed(range(first_block, len(self.enc_blocks))):
            self._block_bwd(self.enc_blocks[i], ws["enc"][i], cur, curb, nxt, nxtb, Me, B, Lk, st)
            cur, curb, nxt, nxtb = nxt, nxtb, cur, curb
            if bucket_cb is not None and i >= 1:
                bucket_cb(i)  # everything that writes this block's gradients is enqueued (main + side stream)
        if stem and first_block == 0:
            # tokens = cat(cls, patch_embed(kept patches)) + frozen pos
            calls.mae_colsum_rows_f32(_ptr(cur), None, B, 1, Lk, 0, De, self.pg(o(ep + "cls_token")), st)
            calls.mae_gather_rows_bf16(_ptr(cur), None, B, Lk - 1, Lk, 1, De, _ptr(ws["dtok"]), st)
            self._linear_bwd_w(_ptr(ws["dtok"]), De, _ptr(ws["patches"]), pd, B * (Lk - 1),
                               o(ep + "patch_embed.proj.weight"), o(ep + "patch_embed.proj.bias"), st)
        if bucket_cb is not None and first_block == 0:
            bucket_cb(0)
        self._wg_join()
        return cur

    def _decode(self, st: int) -> None:
        ws, d = self.ws, self.dims
        B, L, Lk, Lm, De, Dd = d["B"], d["L"], d["Lk"], d["Lm"], d["De"], d["Dd"]
        o = self._o
        calls.mae_gemm_tn(_ptr(ws["xnf"]), self.pw(o("decoder.decoder_embed.weight")), d["Me"], Dd, De, EPI_BF16,
                          self.pf(o("decoder.decoder_embed.bias")), _ptr(ws["emb"]), None, None, st)
        x0 = ws["dec"][0]["x_in"] if self.dec_blocks else ws["x_dec"]
        n_dec = len(self.dec_blocks)
        fused_ln = self.scatter_ln and n_dec > 0 and Dd <= 256
        if fused_ln:
            w0, s0 = self.dec_blocks[0], ws["dec"][0]
            calls.mae_scatter_decoder_tokens_ln(_ptr(ws["emb"]), self.pf(o("decoder.mask_token")),
                                                self.pf(o("decoder.decoder_pos_embed")), _ptr(ws["idx_keep"]),
                                                _ptr(ws["idx_mask"]), B, L, Lk, Dd, self.pf(w0.n1w), self.pf(w0.n1b), LN_EPS,
                                                _ptr(x0), _ptr(s0["xn1"]), _ptr(s0["mean1"]), _ptr(s0["rstd1"]), st)
        else:
            scatter = calls.mae_scatter_decoder_tokens_grouped if self._use_grouped() else calls.mae_scatter_decoder_tokens
            scatter(_ptr(ws["emb"]), self.pf(o("decoder.mask_token")), self.pf(o("decoder.decoder_pos_embed")),
                    _ptr(ws["idx_keep"]), _ptr(ws["idx_mask"]), B, L, Lk, Dd, _ptr(x0), st)
        tail = self._use_masked_tail()
        for i, w in enumerate(self.dec_blocks):
            x_out = ws["dec"][i + 1]["x_in"] if i + 1 < n_dec else ws["x_dec"]
            self._block_fwd(w, ws["dec"][i], x_out, d["Md"], B, L, st, tail=tail and i == n_dec - 1,
                            skip_norm1=fused_ln and i == 0)
        if tail and self._use_fused_mlp(Dd, d["Mm"]):
            # the fused MLP kernel already left the compact sums in x_dec: plain LayerNorm over Mm rows
            calls.mae_layernorm_fwd(_ptr(ws["x_dec"]), self.pf(o("decoder.decoder_norm.weight")),
                                    self.pf(o("decoder.decoder_norm.bias")), None, 0, 0, d["Mm"], Dd, LN_EPS,
                                    _ptr(ws["xnd"]), _ptr(ws["meand"]), _ptr(ws["rstdd"]), st)
        elif tail:
            # decoder_norm over  x_mid[idx_mask rows] + mlp output ; the summed rows (compact) are kept in x_dec for the backward
            calls.mae_layernorm_fwd_add(_ptr(ws["dec"][-1]["x_mid"]), _ptr(ws["dlt"]),
                                        self.pf(o("decoder.decoder_norm.weight")), self.pf(o("decoder.decoder_norm.bias")),
                                        self._rows_mask(), Lm, L, d["Mm"], Dd, LN_EPS, _ptr(ws["x_dec"]),
                                        _ptr(ws["xnd"]), _ptr(ws["meand"]), _ptr(ws["rstdd"]), st)
        else:
            calls.mae_layernorm_fwd(_ptr(ws["x_dec"]), self.pf(o("decoder.decoder_norm.weight")),
                                    self.pf(o("decoder.decoder_norm.bias")), self._rows_mask(), Lm, L, d["Mm"], Dd, LN_EPS,
                                    _ptr(ws["xnd"]), _ptr(ws["meand"]), _ptr(ws["rstdd"]), st)
        calls.mae_gemm_tn(_ptr(ws["xnd"]), self.pw(o("decoder.decoder_pred.weight")), d["Mm"], d["pd"], Dd, EPI_BF16,
                          self.pf(o("decoder.decoder_pred.bias")), _ptr(ws["pred"]), None, None, st)

    # ---------------- backward ----------------
    def _wg_event(self) -> torch.cuda.Event:
        if self._wg_next_event == len(self._wg_events):
            self._wg_events.append(torch.cuda.Event())
        ev = self._wg_events[self._wg_next_event]
        self._wg_next_event += 1
        return ev

    def _wg_wait(self, buf_ptr: int) -> None:
        """Call before the main stream overwrites ``buf_ptr``: waits for the side-stream wgrad that still reads it."""
        ev = self._wg_pending.pop(buf_ptr, None)
        if ev is not None:
            torch.cuda.current_stream(self.flat.device).wait_event(ev)

    def _wg_join(self) -> None:
        """Main stream waits for every outstanding side-stream wgrad (before the gradients are consumed)."""
        if self._wg_pending:
            main = torch.cuda.current_stream(self.flat.device)
            for ev in self._wg_pending.values():
                main.wait_event(ev)
            self._wg_pending.clear()
        self._wg_next_event = 0

    def _side_begin(self, st: int, extra: bool = False) -> int:
        """Stream for work that only feeds the optimizer: the side stream (ordered behind everything enqueued on the main
        stream so far), or ``st`` itself when the two-stream schedule is off."""
        if not self.wgrad_side_stream or (extra and not self.side_extras):
            self._side_skip = True
            return st
        self._side_skip = False
        dev = self.flat.device
        if self._side is None:
            self._side = torch.cuda.Stream(dev, priority=int(os.environ.get("MAE_WGRAD_PRIO", "0")))
        ready = self._wg_event()
        ready.record(torch.cuda.current_stream(dev))
        self._side.wait_event(ready)
        return self._side.cuda_stream

    def _side_end(self, *buf_ptrs: int) -> None:
        """Marks the buffers the side-stream launch just enqueued still reads (or writes): :meth:`_wg_wait` on one of them
        orders the main stream behind it.  The side stream is in-order, so a later event also covers earlier launches."""
        if not self.wgrad_side_stream or self._side_skip:
            return
        done = self._wg_event()
        done.record(self._side)
        for ptr in buf_ptrs:
            self._wg_pending[ptr] = done

    def _linear_bwd_w(self, dy: int, n_out: int, x: int, n_in: int, M: int, w_off: int, b_off: int, st: int) -> None:
        """dW (n_out, n_in) += dY^T X and db += colsum(dY), bias gradient fused into the same launch.  The wider
        operand rides the MMA M axis (chunks of 128), the narrower one the MMA N axis (chunks of <= 256)."""
        st = self._side_begin(st)
        if n_in <= n_out:
            calls.mae_gemm_wgrad(dy, n_out, x, n_in, M, self.pg(w_off), n_in, 1, self.pg(b_off), 0, st)
        else:
            calls.mae_gemm_wgrad(x, n_in, dy, n_out, M, self.pg(w_off), 1, n_in, self.pg(b_off), 1, st)
        self._side_end(dy)

    def _block_bwd(self, w: _BlockW, s: Dict[str, torch.Tensor], dx: torch.Tensor, dxb: torch.Tensor,
                   dx_out: torch.Tensor, dxb_out: torch.Tensor, M: int, B: int, L: int, st: int, tail: bool = False) -> None:
        """dx/dxb: fp32 / bf16 gradient w.r.t. the block output; writes the gradient w.r.t. its input.  ``tail`` (see
        :meth:`_block_fwd`): dx/dxb hold the Mm masked rows only (compact); norm2's backward scatters them back."""
        ws, d = self.ws, self.dims
        D, H, hd = w.dim, w.heads, w.hd
        Mt = d["Mm"] if tail else M
        # mlp.fc2
        self._linear_bwd_w(_ptr(dxb), D, _ptr(s["h"]), 4 * D, Mt, w.fc2w, w.fc2b, st)
        self._wg_wait(_ptr(ws["dh"]))
        calls.mae_gemm_tn(_ptr(dxb), self.pwt(w.fc2w), Mt, 4 * D, D, EPI_DGELU_BF16, None, _ptr(ws["dh"]), None,
                          _ptr(s["pre"]), st)
        # mlp.fc1 + norm2
        self._linear_bwd_w(_ptr(ws["dh"]), 4 * D, _ptr(s["xn2"]), D, Mt, w.fc1w, w.fc1b, st)
        calls.mae_gemm_tn(_ptr(ws["dh"]), self.pwt(w.fc1w), Mt, D, 4 * D, EPI_BF16, None, _ptr(ws["g"]), None, None, st)
        self._wg_wait(_ptr(dxb_out))
        self._wg_wait(_ptr(dx_out))
        if tail:
            # kept rows were zeroed by backward() (side stream); masked rows: LN'(g) + the (compact) residual gradient
            calls.mae_layernorm_bwd_scatter(_ptr(ws["g"]), _ptr(s["x_mid"]), _ptr(s["mean2"]), _ptr(s["rstd2"]),
                                            self.pf(w.n2w), _ptr(dx), self._rows_mask(), d["Lm"], L, Mt, D, _ptr(dx_out),
                                            _ptr(dxb_out), self.pg(w.n2w), self.pg(w.n2b), st)
        else:
            calls.mae_layernorm_bwd(_ptr(ws["g"]), _ptr(s["x_mid"]), _ptr(s["mean2"]), _ptr(s["rstd2"]), self.pf(w.n2w),
                                    _ptr(dx), None, 0, 0, M, D, _ptr(dx_out), _ptr(dxb_out), self.pg(w.n2w), self.pg(w.n2b), st)
        # attn.proj
        self._linear_bwd_w(_ptr(dxb_out), D, _ptr(s["att"]), D, M, w.projw, w.projb, st)
        calls.mae_gemm_tn(_ptr(dxb_out), self.pwt(w.projw), M, D, D, EPI_BF16, None, _ptr(ws["datt"]), None, None, st)
        self._wg_wait(_ptr(ws["dqkv"]))
        if tail and self._use_grouped():  # the gradient of the attention output is zero on the kept tokens' rows
            calls.mae_attention_bwd_q(_ptr(s["qkv"]), _ptr(s["att"]), _ptr(ws["datt"]), _ptr(s["lse"]), B, L, H, hd,
                                      d["Lm"], _ptr(ws["dqkv"]), st)
        else:
            calls.mae_attention_bwd(_ptr(s["qkv"]), _ptr(s["att"]), _ptr(ws["datt"]), _ptr(s["lse"]), B, L, H, hd,
                                    _ptr(ws["dqkv"]), st)
        # attn.qkv + norm1
        self._linear_bwd_w(_ptr(ws["dqkv"]), 3 * D, _ptr(s["xn1"]), D, M, w.qkvw, w.qkvb, st)
        calls.mae_gemm_tn(_ptr(ws["dqkv"]), self.pwt(w.qkvw), M, D, 3 * D, EPI_BF16, None, _ptr(ws["g"]), None, None, st)
        self._wg_wait(_ptr(dxb_out))
        calls.mae_layernorm_bwd(_ptr(ws["g"]), _ptr(s["x_in"]), _ptr(s["mean1"]), _ptr(s["rstd1"]), self.pf(w.n1w),
                                _ptr(dx_out), None, 0, 0, M, D, _ptr(dx_out), _ptr(dxb_out), self.pg(w.n1w),
                                self.pg(w.n1b), st)

    def backward(self, dpred: Optional[torch.Tensor] = None, bucket_cb=None) -> None:
        """Gradients of all trainable parameters into ``self.grads`` (zeroed first).  ``dpred`` defaults to
        the MSE gradient produced by the fused loss kernel in :meth:`forward`.  ``bucket_cb(i)`` fires, from the last
        bucket of :meth:`grad_bucket_boundaries` to the first, as soon as every launch that writes bucket ``i`` has been
        enqueued -- on the main stream or on the weight-gradient side stream (``self._side``): the data-parallel trainer
        starts the bucket's all-reduce behind both, so all but the stem + first-block bucket overlap the rest of backward."""
        ws, d = self.ws, self.dims
        B, L, Lk, Lm, De, Dd, pd = d["B"], d["L"], d["Lk"], d["Lm"], d["De"], d["Dd"], d["pd"]
        Me, Md, Mm = d["Me"], d["Md"], d["Mm"]
        o = self._o
        dev = self.flat.device
        st = torch.cuda.current_stream(dev).cuda_stream
        if dpred is None:
            dpred = ws["dpred"]
        else:
            dpred = dpred.to(torch.bfloat16).contiguous()
        self._wg_wait(self.flat_bf16_t.data_ptr())  # W^T refresh launched by clip_adamw on the side stream
        self.grads.zero_()
        dx_a, dx_b, dxb_a, dxb_b = ws["dx_a"], ws["dx_b"], ws["dxb_a"], ws["dxb_b"]
        # decoder_pred + decoder_norm (gradient lands only on the masked rows of the decoder stream)
        self._linear_bwd_w(_ptr(dpred), pd, _ptr(ws["xnd"]), Dd, Mm, o("decoder.decoder_pred.weight"),
                           o("decoder.decoder_pred.bias"), st)
        calls.mae_gemm_tn(_ptr(dpred), self.pwt(o("decoder.decoder_pred.weight")), Mm, Dd, pd, EPI_BF16, None,
                          _ptr(ws["g"]), None, None, st)
        tail = self._use_masked_tail()
        if tail:
            # kept rows of the last block's mid-stream gradient are zero (no gradient through its MLP): cleared early, off
            # the critical path; norm2's backward (which fills the masked rows) waits for it through dxb_b
            st2 = self._side_begin(st, extra=True)
            calls.mae_zero_rows(_ptr(dx_b), _ptr(dxb_b), self._rows_keep(), B, Lk, L, Dd, st2)
            self._side_end(_ptr(dxb_b))
            # decoder_norm saw the compact masked rows (x_dec holds their pre-norm values): compact gradient
            calls.mae_layernorm_bwd(_ptr(ws["g"]), _ptr(ws["x_dec"]), _ptr(ws["meand"]), _ptr(ws["rstdd"]),
                                    self.pf(o("decoder.decoder_norm.weight")), None, None, 0, 0, Mm, Dd,
                                    _ptr(dx_a), _ptr(dxb_a), self.pg(o("decoder.decoder_norm.weight")),
                                    self.pg(o("decoder.decoder_norm.bias")), st)
        else:
            # only the masked rows receive a gradient from the prediction head: zero the kept rows, LN-bwd writes the rest
            calls.mae_zero_rows(_ptr(dx_a), _ptr(dxb_a), self._rows_keep(), B, Lk, L, Dd, st)
            calls.mae_layernorm_bwd(_ptr(ws["g"]), _ptr(ws["x_dec"]), _ptr(ws["meand"]), _ptr(ws["rstdd"]),
                                    self.pf(o("decoder.decoder_norm.weight")), None, self._rows_mask(), Lm, L, Mm, Dd,
                                    _ptr(dx_a), _ptr(dxb_a), self.pg(o("decoder.decoder_norm.weight")),
                                    self.pg(o("decoder.decoder_norm.bias")), st)
        cur, curb, nxt, nxtb = dx_a, dxb_a, dx_b, dxb_b
        n_dec = len(self.dec_blocks)
        n_enc = len(self.enc_blocks)
        for i in reversed(range(n_dec)):
            self._block_bwd(self.dec_blocks[i], ws["dec"][i], cur, curb, nxt, nxtb, Md, B, L, st,
                            tail=tail and i == n_dec - 1)
            cur, curb, nxt, nxtb = nxt, nxtb, cur, curb
            if bucket_cb is not None and i >= 1:
                bucket_cb(n_enc + i)
        # decoder input = scatter(embed) / mask_token, + frozen pos
        st2 = self._side_begin(st, extra=True)
        calls.mae_colsum_rows_f32(_ptr(cur), self._rows_mask(), B, Lm, L, 0, Dd, self.pg(o("decoder.mask_token")), st2)
        self._side_end(_ptr(cur))
        calls.mae_gather_rows_bf16(_ptr(cur), self._rows_keep(), B, Lk, L, 0, Dd, _ptr(ws["demb"]), st)
        self._linear_bwd_w(_ptr(ws["demb"]), Dd, _ptr(ws["xnf"]), De, Me, o("decoder.decoder_embed.weight"),
                           o("decoder.decoder_embed.bias"), st)
        if bucket_cb is not None:
            bucket_cb(n_enc)  # mask token + decoder_embed + decoder block 0
        calls.mae_gemm_tn(_ptr(ws["demb"]), self.pwt(o("decoder.decoder_embed.weight")), Me, De, Dd, EPI_BF16, None,
                          _ptr(ws["g"]), None, None, st)
        self._encoder_bwd(_ptr(ws["g"]), cur, curb, nxt, nxtb, st, bucket_cb=bucket_cb)

    # ---------------- optimizer ----------------
    def clip_adamw(self, lr: float, weight_decay: float = 0.05, betas=(0.9, 0.999), eps: float = 1e-8,
                   max_norm: float = 1.0, grads: Optional[torch.Tensor] = None, n_partials: int = 0) -> torch.Tensor:
        """Lightning clip-by-norm + AdamW over the flat buffers; refreshes the bf16 shadows.  Returns the
        (device) pre-clip gradient norm.  ``grads`` / ``n_partials``: an already reduced gradient buffer whose sum-of-
        squares partials are in ``self.partials[:n_partials]`` (the peer-memory all-reduce produces both)."""
        st = torch.cuda.current_stream(self.flat.device).cuda_stream
        self._wg_join()  # (a no-op after backward(); orders a pending W^T refresh before the weights change again)
        self.step_count += 1
        if grads is None:
            grads, nparts = self.grads, 1024
            calls.mae_grad_sumsq(grads.data_ptr(), self.n_opt, self.partials.data_ptr(), nparts, st)
        else:
            nparts = n_partials
        calls.mae_clip_adamw(self.flat.data_ptr(), grads.data_ptr(), self.exp_avg.data_ptr(),
                             self.exp_avg_sq.data_ptr(), self.flat_bf16.data_ptr(), self.n_opt, float(lr), betas[0],
                             betas[1], eps, weight_decay, max_norm, self.step_count, self.partials.data_ptr(), nparts,
                             self.grad_norm.data_ptr(), st)
        # W^T shadows are read by the dgrad GEMMs only: off the next forward's critical path
        st2 = self._side_begin(st, extra=True)
        calls.mae_transpose_cast_bf16(self.flat.data_ptr(), self.flat_bf16_t.data_ptr(), self.t_table.data_ptr(),
                                      self.t_table.shape[0], self.t_max_rows, self.t_max_cols, st2)
        self._side_end(self.flat_bf16_t.data_ptr())
        self._shadow_fresh, self._shadow_version = True, self._param_versions()
        return self.grad_norm

    def named_grads(self) -> Dict[str, torch.Tensor]:
        out = {}
        for n, p in self._ordered_params():
            o = self.offsets[n]
            if o < self.n_opt:
                out[n] = self.grads[o:o + p.numel()].view(p.shape)
        return out


class _MAEFunction(torch.autograd.Function):
    """autograd bridge for the reference-style API: ``pred, target = model(images)`` followed by any loss."""

    @staticmethod
    def forward(ctx, engine: MAEEngine, images, noise, mask_ratio, *params):
        pred, target, _ = engine.forward(images, noise, mask_ratio, want_target=True, compute_loss=False)
        ctx.engine = engine
        ctx.n_params = len(params)
        ctx.mark_non_differentiable(target)
        return pred.clone(), target

    @staticmethod
    def backward(ctx, dpred, _dtarget):
        eng: MAEEngine = ctx.engine
        eng.backward(dpred)
        grads = eng.named_grads()
        out = []
        for n, p in eng._grad_order:
            g = grads.get(n)
            out.append(g.clone() if (g is not None and p.requires_grad) else None)
        return (None, None, None, None, *out)


class MaskedAutoencoder(nn.Module):
    """Masked Autoencoder (MAE) with ViT backbone -- drop-in for reference ``src/models/mae.py``."""

    def __init__(self, general_cfg: Dict[str, Any], encoder_cfg: Dict[str, Any], decoder_cfg: Dict[str, Any]):
        super().__init__()
        self.mask_ratio = general_cfg.get("mask_ratio", 0.75)
        self.image_size = general_cfg.get("image_size", 96)
        self.patch_size = general_cfg.get("patch_size", 6)
        self.in_chans = general_cfg.get("in_chans", 3)
        embed_dim = encoder_cfg.get("embed_dim", 384)
        vit = _VisionTransformer(self.image_size, self.patch_size, self.in_chans, embed_dim,
                                 encoder_cfg.get("depth", 12), encoder_cfg.get("num_heads", 6))
        self.encoder = _MaskedViT(vit)
        self.sequence_length = self.encoder.sequence_length
        self.decoder = _MAEDecoder(vit.patch_embed.num_patches, self.patch_size, self.in_chans, embed_dim,
                                   decoder_cfg.get("decoder_embed_dim", 512), decoder_cfg.get("decoder_depth", 4),
                                   decoder_cfg.get("decoder_num_heads", 6))
        object.__setattr__(self, "_engine", MAEEngine(self))
        self._check_supported()

    def _check_supported(self) -> None:
        vit, dec = self.encoder.vit, self.decoder
        pd = self.in_chans * self.patch_size**2
        dims = [(vit.embed_dim, vit.num_heads)]
        if len(dec.decoder_blocks):
            dims.append((dec.decoder_embed.out_features, dec.decoder_blocks[0].attn.num_heads))
        for D, H in dims:
            if D % 16 or D > 1024 or D // H not in (24, 32, 64):
                raise NotImplementedError(
                    f"kernels cover embed dims <= 1024 (multiple of 16) with head_dim 24/32/64; got {D}/{H}")
        if pd % 16 or pd > 768 or self.sequence_length > 208:
            raise NotImplementedError("kernels cover patch_dim % 16 == 0, patch_dim <= 768 and sequence_length <= 208")

    @property
    def engine(self) -> MAEEngine:
        return self._engine

    # -- reference surface ---------------------------------------------------------------
    def forward(self, images: torch.Tensor, *, mask_ratio: Optional[float] = None, noise: Optional[torch.Tensor] = None):
        """``(x_pred (B,Lm,p*p*C) bf16, target (B,Lm,p*p*C) f32)`` exactly like reference ``forward``."""
        eng = self._engine
        eng._ensure_flat(images.device)
        eng._grad_order = [(n, p) for n, p in self.named_parameters()]
        params = [p for _, p in eng._grad_order]
        r = self.mask_ratio if mask_ratio is None else mask_ratio
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            return _MAEFunction.apply(eng, images, noise, r, *params)
        pred, target, _ = eng.forward(images, noise, r, want_target=True, compute_loss=False)
        return pred.clone(), target

    def forward_loss(self, images: torch.Tensor, mask_ratio: Optional[float] = None,
                     noise: Optional[torch.Tensor] = None):
        """Fused path: ``(loss (), pred, mask (B, num_patches) with 1 = masked)``; follow with
        ``engine.backward()`` / ``engine.clip_adamw()``.  The loss stays on the device."""
        r = self.mask_ratio if mask_ratio is None else mask_ratio
        pred, _, loss = self._engine.forward(images, noise, r)
        ws = self._engine.ws
        mask = torch.ones(images.shape[0], self.sequence_length, device=images.device)
        mask.scatter_(1, ws["idx_keep"], 0.0)
        return loss[0].clone(), pred, mask[:, 1:]

    @torch.no_grad()
    def forward_encoder(self, images: torch.Tensor, idx_keep=None):
        """``encoder.encode(images, idx_keep)``: LayerNorm-ed encoder tokens (B, num_keep, D) (inference only)."""
        eng = self._engine
        eng._ensure_flat(images.device)
        B, L = images.shape[0], self.sequence_length
        if idx_keep is None:
            idx_keep = torch.arange(L, device=images.device).unsqueeze(0).expand(B, -1)
        idx_keep = idx_keep.to(torch.int64).contiguous()
        if idx_keep.shape[1] > 208:
            raise NotImplementedError("sequence too long for the fused attention kernel")
        eng._ensure_ws(B, idx_keep.shape[1], images.device)
        st = torch.cuda.current_stream(images.device).cuda_stream
        eng._refresh_shadow(st)
        eng.ws["idx_keep"].copy_(idx_keep)
        eng._encode(images.contiguous(), eng.ws["idx_keep"], st)
        return eng.ws["xnf"].view(B, idx_keep.shape[1], -1).float()

    @torch.no_grad()
    def forward_decoder(self, x_encoded, idx_keep, idx_mask):
        """``decoder.embed`` -> mask-token scatter -> ``decode`` -> gather ``idx_mask`` -> ``predict``."""
        eng = self._engine
        B = x_encoded.shape[0]
        eng._ensure_flat(x_encoded.device)
        eng._ensure_ws(B, idx_keep.shape[1], x_encoded.device)
        st = torch.cuda.current_stream(x_encoded.device).cuda_stream
        eng._refresh_shadow(st)
        eng.ws["idx_keep"].copy_(idx_keep)
        eng.ws["idx_mask"].copy_(idx_mask)
        eng.ws["xnf"].copy_(x_encoded.reshape(eng.ws["xnf"].shape))
        eng._decode(st)
        return eng.ws["pred"].clone()
