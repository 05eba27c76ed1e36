"""CPU oracle for the MAE pre-training hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module.  The product path
(``ssrl-vit-mae-jepa_b200``) never imports it and has no CPU fallback.

What this restates
------------------
A plain fp32 PyTorch restatement of the reference's MAE step
(``/root/reference/src/models/mae.py:15-94`` + ``src/training/mae.py:40-76``) including
the arithmetic that lives in the reference's third-party dependencies, which are
absent from ``/root/reference`` and are not installed in this image:

* ``lightly==1.5.22`` (reference ``uv.lock:866-867``): ``lightly.models.utils``
  {random_token_mask, get_at_index, set_at_index, repeat_token, patchify},
  ``MaskedVisionTransformerTIMM``, ``MAEDecoderTIMM`` -- called at
  ``src/models/mae.py:38,45-52,55,59,61-73,79-83,90-92``.
* ``timm==1.0.21`` (``uv.lock:2163-2164``): ``VisionTransformer``/``Block``/``Attention``/
  ``Mlp``/``PatchEmbed`` -- built at ``src/models/mae.py:28-36``.

Their published algorithms are restated here; nothing is copied.

Pinning status
--------------
The reference's own tests hold NO golden vectors for this path (its ``tests/`` are two
CUDA install smoke scripts) and the reference cannot be imported here, so by the
reference's own artefacts this oracle is "parity unpinned".  What does pin it
(``tests/test_oracle.py``):

1. the recorded Lightning model summaries in ``notebook.ipynb:986-995`` (2.0 M trainable /
   48.7 K frozen / 152 modules) and ``notebook.ipynb:1689-1698`` (bare ViT: 103 modules,
   1,051,920 params) -> exact integers 2,035,104 / 1,986,384 / 48,720 / 151+1 / 100;
2. an INDEPENDENT implementation present in the image, HuggingFace
   ``transformers.models.vit_mae`` (``ViTMAEForPreTraining``): with weights mapped across
   and identical noise, loss / predictions / sin-cos tables / patchify agree to fp32
   round-off (``tests/golden/make_golden.py`` asserts this when it writes the fixtures);
3. shape/identity invariants the reference's callers rely on
   (``scripts/evaluation/visualize_reconstruction.py:192-231``).

Numerics: fp32, ``torch.set_float32_matmul_precision("highest")``, stable
argsort.
"""
from __future__ import annotations

import math
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

# --------------------------------------------------------------------------------------
# lightly.models.utils restatements
# --------------------------------------------------------------------------------------


def num_keep_tokens(seq_len: int, mask_ratio: float) -> int:
    """lightly ``random_token_mask``: ``int(L*(1-mask_ratio))`` then ``max(1, .)``."""
    return max(1, int(seq_len * (1 - mask_ratio)))


def random_token_mask_from_noise(noise: torch.Tensor, mask_ratio: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """``utils.random_token_mask`` (called at reference ``src/models/mae.py:79-83``) with the
    ``torch.rand`` draw hoisted out so tests can inject it.  ``noise`` is (B, L) fp32; column 0
    (class token) is forced to -1 so it always sorts first.  Ties break towards the lower index
    (stable sort).  Returns ``(idx_keep (B,num_keep), idx_mask (B,L-num_keep))`` int64.
    """
    B, L = noise.shape
    k = num_keep_tokens(L, mask_ratio)
    noise = noise.clone()
    noise[:, 0] = -1.0
    idx = torch.argsort(noise, dim=1, stable=True)
    return idx[:, :k].contiguous(), idx[:, k:].contiguous()


def random_token_mask_numpy(noise: np.ndarray, mask_ratio: float) -> Tuple[np.ndarray, np.ndarray]:
    """numpy twin of :func:`random_token_mask_from_noise` (integer path, bit-exact check)."""
    B, L = noise.shape
    k = num_keep_tokens(L, mask_ratio)
    n = noise.astype(np.float32).copy()
    n[:, 0] = -1.0
    idx = np.argsort(n, axis=1, kind="stable").astype(np.int64)
    return idx[:, :k].copy(), idx[:, k:].copy()


def get_at_index(tokens: torch.Tensor, index: torch.Tensor) -> torch.Tensor:
    """``utils.get_at_index``: gather rows along dim 1 (reference ``src/models/mae.py:72,92``)."""
    return torch.gather(tokens, 1, index.unsqueeze(-1).expand(-1, -1, tokens.shape[-1]))


def set_at_index(tokens: torch.Tensor, index: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    """``utils.set_at_index``: out-of-place scatter along dim 1 (``src/models/mae.py:64-69``)."""
    return torch.scatter(tokens, 1, index.unsqueeze(-1).expand(-1, -1, tokens.shape[-1]), value)


def repeat_token(token: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
    """``utils.repeat_token``: (1,1,D) -> (B,L,D) (``src/models/mae.py:61-64``)."""
    return token.repeat(size[0], size[1], 1)


def patchify(images: torch.Tensor, patch_size: int) -> torch.Tensor:
    """``utils.patchify`` (``src/models/mae.py:90``): (N,C,H,W) -> (N,(H/p)^2,p*p*C); element order
    inside a patch is (py, px, c); patch index = row*(W/p)+col."""
    N, C, H, W = images.shape
    assert H == W and H % patch_size == 0
    g = H // patch_size
    x = images.reshape(N, C, g, patch_size, g, patch_size)
    x = torch.einsum("nchpwq->nhwpqc", x)
    return x.reshape(N, g * g, patch_size * patch_size * C)


def sincos_2d_pos_embed(embed_dim: int, grid_size: int, cls_token: bool = True) -> torch.Tensor:
    """Frozen 2-D sin-cos table used for ``encoder.vit.pos_embed`` and ``decoder.decoder_pos_embed``
    (SURVEY §8a note).  First D/2 channels encode the column (w), last D/2 the row (h); each half is
    ``[sin(pos*w_i), cos(pos*w_i)]`` with ``w_i = 1/10000**(i/(D/4))``; the CLS row is zeros."""
    assert embed_dim % 4 == 0
    gh = np.arange(grid_size, dtype=np.float64)
    gw = np.arange(grid_size, dtype=np.float64)
    grid_w, grid_h = np.meshgrid(gw, gh)  # w varies fastest

    def one_dim(d: int, pos: np.ndarray) -> np.ndarray:
        omega = np.arange(d // 2, dtype=np.float64) / (d / 2.0)
        omega = 1.0 / 10000**omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    emb = np.concatenate([one_dim(embed_dim // 2, grid_w), one_dim(embed_dim // 2, grid_h)], axis=1)
    if cls_token:
        emb = np.concatenate([np.zeros((1, embed_dim)), emb], axis=0)
    return torch.from_numpy(emb).float().unsqueeze(0)


# --------------------------------------------------------------------------------------
# timm 1.0.21 VisionTransformer restatement.  Placeholder Identity /
# Dropout modules are kept so the module count matches the notebook's model summary pins.
# --------------------------------------------------------------------------------------


class PatchEmbed(nn.Module):
    def __init__(self, img_size: int, patch_size: int, in_chans: int, embed_dim: int):
        super().__init__()
        self.img_size = img_size
        self.patch_size = patch_size
        self.grid_size = img_size // patch_size
        self.num_patches = self.grid_size**2
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=True)
        self.norm = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        assert x.shape[-1] == self.img_size and x.shape[-2] == self.img_size
        x = self.proj(x).flatten(2).transpose(1, 2)
        return self.norm(x)


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int):
        super().__init__()
        assert dim % num_heads == 0
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim**-0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.q_norm = nn.Identity()
        self.k_norm = nn.Identity()
        self.attn_drop = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(0.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, N, C = x.shape
        qkv = self.qkv(x).reshape(B, N, 3, self.num_heads, self.head_dim).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        x = F.scaled_dot_product_attention(q, k, v)
        x = x.transpose(1, 2).reshape(B, N, C)
        return self.proj_drop(self.proj(self.norm(x)))


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.drop1 = nn.Dropout(0.0)
        self.norm = nn.Identity()
        self.fc2 = nn.Linear(hidden, dim)
        self.drop2 = nn.Dropout(0.0)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.drop2(self.fc2(self.norm(self.drop1(self.act(self.fc1(x))))))


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, num_heads)
        self.ls1 = nn.Identity()
        self.drop_path1 = nn.Identity()
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.ls2 = nn.Identity()
        self.drop_path2 = nn.Identity()

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x = x + self.drop_path1(self.ls1(self.attn(self.norm1(x))))
        x = x + self.drop_path2(self.ls2(self.mlp(self.norm2(x))))
        return x


class VisionTransformer(nn.Module):
    """timm ``VisionTransformer(img_size, patch_size, in_chans, embed_dim, depth, num_heads,
    num_classes=0)`` as built at reference ``src/models/mae.py:28-36``."""

    def __init__(self, img_size=96, patch_size=8, in_chans=3, embed_dim=144, depth=4, num_heads=6, num_classes=0):
        super().__init__()
        assert num_classes == 0
        self.embed_dim = embed_dim
        self.num_prefix_tokens = 1
        self.patch_embed = PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, embed_dim) * 0.02)
        self.pos_drop = nn.Dropout(0.0)
        self.patch_drop = nn.Identity()
        self.norm_pre = nn.Identity()
        self.blocks = nn.Sequential(*[Block(embed_dim, num_heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim, eps=1e-6)
        self.fc_norm = nn.Identity()
        self.head_drop = nn.Dropout(0.0)
        self.head = nn.Identity()

    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        x = self.patch_embed(x)
        x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1)
        x = self.pos_drop(x + self.pos_embed)
        x = self.norm_pre(self.patch_drop(x))
        return self.norm(self.blocks(x))


def _init_linear_ln(module: nn.Module) -> None:
    for m in module.modules():
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.LayerNorm):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)


class MaskedVisionTransformerTIMM(nn.Module):
    """lightly ``MaskedVisionTransformerTIMM(vit)``, used at ``src/models/mae.py:38``."""

    def __init__(self, vit: VisionTransformer):
        super().__init__()
        self.vit = vit
        self.mask_token = nn.Parameter(torch.zeros(1, 1, vit.embed_dim))
        w = self.vit.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.normal_(self.vit.cls_token, std=0.02)
        _init_linear_ln(self.vit)
        self.vit.pos_embed.data.copy_(sincos_2d_pos_embed(vit.embed_dim, vit.patch_embed.grid_size))
        self.vit.pos_embed.requires_grad = False

    @property
    def sequence_length(self) -> int:
        return self.vit.patch_embed.num_patches + self.vit.num_prefix_tokens

    def preprocess(self, images: torch.Tensor) -> torch.Tensor:
        t = self.vit.patch_embed(images)
        t = torch.cat([self.vit.cls_token.expand(t.shape[0], -1, -1), t], dim=1)
        return self.vit.pos_drop(t + self.vit.pos_embed)

    def encode(self, images: torch.Tensor, idx_keep: Optional[torch.Tensor] = None) -> torch.Tensor:
        t = self.preprocess(images)
        if idx_keep is not None:
            t = get_at_index(t, idx_keep)
        t = self.vit.norm_pre(t)
        t = self.vit.blocks(t)
        return self.vit.norm(t)


class MAEDecoderTIMM(nn.Module):
    """lightly ``MAEDecoderTIMM``, built at ``src/models/mae.py:45-52``."""

    def __init__(self, num_patches, patch_size, embed_dim, decoder_embed_dim, decoder_depth, decoder_num_heads,
                 in_chans=3, mlp_ratio=4.0):
        super().__init__()
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, num_patches + 1, decoder_embed_dim), requires_grad=False)
        self.decoder_blocks = nn.Sequential(
            *[Block(decoder_embed_dim, decoder_num_heads, mlp_ratio) for _ in range(decoder_depth)])
        self.decoder_norm = nn.LayerNorm(decoder_embed_dim, eps=1e-6)
        self.decoder_pred = nn.Linear(decoder_embed_dim, patch_size**2 * in_chans, bias=True)
        nn.init.normal_(self.mask_token, std=0.02)
        self.decoder_pos_embed.data.copy_(sincos_2d_pos_embed(decoder_embed_dim, int(round(math.sqrt(num_patches)))))
        _init_linear_ln(self)

    def embed(self, x):
        return self.decoder_embed(x)

    def decode(self, x):
        return self.decoder_norm(self.decoder_blocks(x + self.decoder_pos_embed))

    def predict(self, x):
        return self.decoder_pred(x)


class MaskedAutoencoder(nn.Module):
    """Restatement of reference ``src/models/mae.py:12-94`` with ``noise`` injectable."""

    def __init__(self, general_cfg: Dict[str, Any], encoder_cfg: Dict[str, Any], decoder_cfg: Dict[str, Any]):
        super().__init__()
        self.mask_ratio = general_cfg.get("mask_ratio", 0.75)
        self.image_size = general_cfg.get("image_size", 96)
        self.patch_size = general_cfg.get("patch_size", 6)
        self.in_chans = general_cfg.get("in_chans", 3)
        vit = VisionTransformer(
            img_size=self.image_size, patch_size=self.patch_size, in_chans=self.in_chans,
            embed_dim=encoder_cfg.get("embed_dim", 384), depth=encoder_cfg.get("depth", 12),
            num_heads=encoder_cfg.get("num_heads", 6), num_classes=0)
        self.encoder = MaskedVisionTransformerTIMM(vit=vit)
        self.sequence_length = self.encoder.sequence_length
        self.decoder = MAEDecoderTIMM(
            num_patches=vit.patch_embed.num_patches, patch_size=self.patch_size, in_chans=self.in_chans,
            embed_dim=encoder_cfg.get("embed_dim", 384),
            decoder_embed_dim=decoder_cfg.get("decoder_embed_dim", 512),
            decoder_depth=decoder_cfg.get("decoder_depth", 4),
            decoder_num_heads=decoder_cfg.get("decoder_num_heads", 6))

    def forward_encoder(self, images, idx_keep=None):
        return self.encoder.encode(images=images, idx_keep=idx_keep)

    def forward_decoder(self, x_encoded, idx_keep, idx_mask):
        B = x_encoded.shape[0]
        x_decode = self.decoder.embed(x_encoded)
        x_masked = repeat_token(self.decoder.mask_token, (B, self.sequence_length))
        x_masked = set_at_index(x_masked, idx_keep, x_decode.type_as(x_masked))
        x_decoded = self.decoder.decode(x_masked)
        x_pred = get_at_index(x_decoded, idx_mask)
        return self.decoder.predict(x_pred)

    def forward(self, images: torch.Tensor, noise: Optional[torch.Tensor] = None):
        B = images.shape[0]
        if noise is None:
            noise = torch.rand(B, self.sequence_length, device=images.device)
        idx_keep, idx_mask = random_token_mask_from_noise(noise, self.mask_ratio)
        x_encoded = self.forward_encoder(images, idx_keep)
        x_pred = self.forward_decoder(x_encoded, idx_keep, idx_mask)
        patches = patchify(images, self.patch_size)
        target = get_at_index(patches, torch.clamp(idx_mask - 1, min=0))
        return x_pred, target


# --------------------------------------------------------------------------------------
# Training-step restatement (reference src/training/mae.py:45-76 + Lightning clip, App. A.7)
# --------------------------------------------------------------------------------------

MAE_YAML = {  # configs/mae.yaml:1-16,21-31 (values only)
    "general": {"image_size": 96, "patch_size": 8, "in_chans": 3},
    "encoder": {"embed_dim": 144, "depth": 4, "num_heads": 6},
    "decoder": {"decoder_embed_dim": 192, "decoder_depth": 2, "decoder_num_heads": 6},
}
MAE_YAML_PRETRAIN = {"mask_ratio_start": 0.75, "mask_ratio_end": 0.75, "mask_ramp_epochs": 5, "total_epochs": 800,
                     "warmup_epochs": 20, "batch_size": 2000, "base_learning_rate": 1.5e-4, "weight_decay": 0.05}


def build_mae(seed: int = 73, mask_ratio: float = 0.75) -> MaskedAutoencoder:
    torch.manual_seed(seed)
    g = dict(MAE_YAML["general"])
    g["mask_ratio"] = mask_ratio
    return MaskedAutoencoder(g, MAE_YAML["encoder"], MAE_YAML["decoder"])


def mask_ratio_at_epoch(epoch: int, start: float, end: float, ramp_epochs: int) -> float:
    """reference ``src/training/mae.py:78-82``."""
    progress = min(epoch / max(1, ramp_epochs - 1), 1.0)
    return start + progress * (end - start)


def lr_lambda(epoch: int, warmup_epochs: int, total_epochs: int) -> float:
    """reference ``src/training/mae.py:67-70``."""
    warmup = (epoch + 1) / max(1, warmup_epochs)
    cosine = 0.5 * (1 + math.cos(math.pi * epoch / total_epochs))
    return min(warmup, 1.0) * cosine


def loss_and_grads(model: MaskedAutoencoder, images: torch.Tensor, noise: torch.Tensor):
    """One fp32 forward + MSELoss + backward (``src/training/mae.py:45-50``).  Returns
    ``(loss, pred, target, {name: grad})`` with grads detached; parameters' ``.grad`` is left set."""
    for p in model.parameters():
        p.grad = None
    pred, target = model(images, noise=noise)
    loss = F.mse_loss(pred, target)
    loss.backward()
    grads = {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}
    return loss.detach(), pred.detach(), target.detach(), grads


def clip_adamw_reference(params, grads, m, v, step: int, lr: float, wd: float = 0.05, betas=(0.9, 0.999),
                         eps: float = 1e-8, max_norm: float = 1.0):
    """Lightning ``gradient_clip_val=1.0`` (norm) + ``torch.optim.AdamW`` single-tensor math (App. A.7),
    written out explicitly (numpy-style, in place) so the fused kernel has an independent check."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    b1, b2 = betas
    bc1 = 1 - b1**step
    bc2 = 1 - b2**step
    for p, g, mi, vi in zip(params, grads, m, v):
        g = g * coef
        p.mul_(1 - lr * wd)
        mi.mul_(b1).add_(g, alpha=1 - b1)
        vi.mul_(b2).addcmul_(g, g, value=1 - b2)
        denom = (vi.sqrt() / math.sqrt(bc2)).add_(eps)
        p.addcdiv_(mi, denom, value=-lr / bc1)
    return total


class OracleTrainer:
    """fwd + MSE + bwd + clip(1.0) + AdamW on CPU: the timed ``cpu_baseline`` step and the
    multi-step parity reference."""

    def __init__(self, model: MaskedAutoencoder, batch_size: int, base_lr: float = 1.5e-4, wd: float = 0.05,
                 warmup_epochs: int = 20, total_epochs: int = 800):
        self.model = model
        self.lr = base_lr * batch_size / 256  # src/training/mae.py:60
        self.opt = torch.optim.AdamW(model.parameters(), lr=self.lr, weight_decay=wd)
        self.sched = torch.optim.lr_scheduler.LambdaLR(
            self.opt, lambda e: lr_lambda(e, warmup_epochs, total_epochs))

    def step(self, images: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.opt.zero_grad(set_to_none=True)
        pred, target = self.model(images, noise=noise)
        loss = F.mse_loss(pred, target)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.model.parameters(), 1.0)
        self.opt.step()
        return loss.detach()


# --------------------------------------------------------------------------------------
# Downstream classifier restatement (reference src/models/classifier.py:10-57)
# --------------------------------------------------------------------------------------
class ViTClassifier(nn.Module):
    """``encoder.forward_features`` -> CLS / mean pooling -> ``head.classification`` Linear."""

    def __init__(self, encoder: VisionTransformer, num_classes: int = 10, pool: str = "cls"):
        super().__init__()
        self.encoder = encoder
        self.pool_type = pool
        self.head = nn.Module()
        self.head.classification = nn.Linear(encoder.embed_dim, num_classes)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        feats = self.encoder.forward_features(x)
        pooled = feats[:, 0] if self.pool_type == "cls" else feats.mean(dim=1)
        return self.head.classification(pooled)
