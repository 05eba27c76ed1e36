"""Downstream classifier path (BASELINE config #4; reference ``src/models/classifier.py`` and
``src/training/classifier.py``): ``ViTClassifier`` = encoder ``forward_features`` over all 145 tokens -> CLS /
mean pooling -> ``Linear(embed_dim, num_classes)``, trained with cross-entropy, with the reference's freeze /
unfreeze-last-N policies.

The encoder (99.9 % of the work) runs through the same sm_100a kernels as the MAE step (``ViTEngine``: patch
gather + tcgen05 GEMMs + fused attention + LayerNorm, forward and backward); the pooling, the 144x10 head and the
cross-entropy on (B, 10) logits are plain torch ops exactly as in the reference.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional

import torch
import torch.nn.functional as F
from torch import nn
from torch.optim import AdamW
from torch.optim.lr_scheduler import LambdaLR

from ._lib import calls
from .mae import MAEEngine, _VisionTransformer, _ptr
from .training import _Base, lr_factor

VisionTransformer = _VisionTransformer  # timm-style constructor, see its docstring


class ViTEngine(MAEEngine):
    """Flat storage + kernel sequencing for a bare ViT (``forward_features`` and its backward)."""

    NOGRAD = ()      # a stand-alone timm ViT trains pos_embed too; its gradient is a batch sum (see backward)
    EP = ""

    def _vit(self):
        return self.model

    def _ensure_ws(self, B: int, L: int, device: torch.device) -> None:
        key = (B, L, device)
        if self._ws_key == key:
            return
        vit = self.model
        De, pd = vit.embed_dim, vit.patch_embed.in_chans * vit.patch_embed.patch_size**2
        Me = B * L
        bf, f32 = torch.bfloat16, torch.float32

        def e(shape, dt):
            return torch.empty(shape, dtype=dt, device=device)

        def block_ws(M, D, H):
            return {"x_in": e((M, D), f32), "xn1": e((M, D), bf), "mean1": e((M,), f32), "rstd1": e((M,), f32),
                    "qkv": e((M, 3 * D), bf), "att": e((M, D), bf), "lse": e((M * H,), f32),
                    "x_mid": e((M, D), f32), "xn2": e((M, D), bf), "mean2": e((M,), f32), "rstd2": e((M,), f32),
                    "pre": e((M, 4 * D), bf), "h": e((M, 4 * D), bf)}

        ws: Dict[str, Any] = {"idx_keep": torch.arange(L, device=device).repeat(B, 1).contiguous(),
                              "patches": e((B * (L - 1), pd), bf), "tok": e((B * (L - 1), De), bf),
                              "enc": [block_ws(Me, De, b.heads) for b in self.enc_blocks], "x_enc": e((Me, De), f32),
                              "xnf": e((Me, De), bf), "meanf": e((Me,), f32), "rstdf": e((Me,), f32),
                              "dx_a": e((Me * De,), f32), "dx_b": e((Me * De,), f32), "dxb_a": e((Me * De,), bf),
                              "dxb_b": e((Me * De,), bf), "dh": e((Me * De * 4,), bf), "g": e((Me * De,), bf),
                              "datt": e((Me * De,), bf), "dqkv": e((Me * De * 3,), bf), "dtok": e((B * (L - 1), De), bf)}
        self.ws, self._ws_key = ws, key
        self.dims = dict(B=B, L=L, Lk=L, Lm=0, De=De, Dd=De, pd=pd, Me=Me, Md=0, Mm=0)

    def features(self, images: torch.Tensor) -> torch.Tensor:
        vit = self.model
        pe = vit.patch_embed
        if not images.is_cuda:
            raise RuntimeError("the ViT runs only on CUDA tensors (sm_100a kernels; no CPU fallback)")
        if images.dtype != torch.float32 or images.dim() != 4 or images.shape[1] != pe.in_chans or \
                images.shape[2] != pe.img_size or images.shape[3] != pe.img_size:
            raise ValueError(f"expected float32 images (B,{pe.in_chans},{pe.img_size},{pe.img_size})")
        images = images.contiguous()
        dev, B, L = images.device, images.shape[0], pe.num_patches + 1
        self._ensure_flat(dev)
        self._ensure_ws(B, L, dev)
        st = torch.cuda.current_stream(dev).cuda_stream
        self._refresh_shadow(st)
        self._encode(images, self.ws["idx_keep"], st)
        return self.ws["xnf"].view(B, L, -1).float()

    def features_backward(self, dfeats: torch.Tensor, first_block: int, stem: bool) -> Optional[torch.Tensor]:
        """Parameter gradients of blocks[first_block:], the final norm and (if ``stem``) cls_token / patch-embed into
        ``self.grads``; returns d(loss)/d(token embeddings) (B, L, D) when ``stem`` (for a trainable pos_embed)."""
        ws, d = self.ws, self.dims
        st = torch.cuda.current_stream(self.flat.device).cuda_stream
        self.grads.zero_()
        g = dfeats.reshape(d["Me"], d["De"]).to(torch.bfloat16).contiguous()
        cur = self._encoder_bwd(_ptr(g), ws["dx_a"], ws["dxb_a"], ws["dx_b"], ws["dxb_b"], st, first_block, stem)
        return cur[:d["Me"] * d["De"]].view(d["B"], d["L"], d["De"]) if (stem and first_block == 0) else None


class _ViTFeatures(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine: ViTEngine, images, names, *params):
        ctx.engine, ctx.names, ctx.req = engine, names, [p.requires_grad for p in params]
        return engine.features(images)

    @staticmethod
    def backward(ctx, dfeats):
        eng: ViTEngine = ctx.engine
        trainable = {n for n, r in zip(ctx.names, ctx.req) if r}
        blocks = [int(n.split(".")[1]) for n in trainable if n.startswith("blocks.")]
        stem = any(n in trainable for n in ("cls_token", "pos_embed", "patch_embed.proj.weight", "patch_embed.proj.bias"))
        first_block = 0 if stem else (min(blocks) if blocks else len(eng.enc_blocks))
        dtok = eng.features_backward(dfeats, first_block, stem)
        grads = eng.named_grads()
        out = []
        for n, r in zip(ctx.names, ctx.req):
            if not r:
                out.append(None)
            elif n == "pos_embed":
                out.append(dtok.sum(0, keepdim=True))
            else:
                out.append(grads[n].clone())
        return (None, None, None, *out)


def vit_forward_features(vit: _VisionTransformer, x: torch.Tensor) -> torch.Tensor:
    eng = vit.__dict__.get("_vit_engine")
    if eng is None:
        eng = ViTEngine(vit)
        object.__setattr__(vit, "_vit_engine", eng)
    named = list(vit.named_parameters())
    if torch.is_grad_enabled() and any(p.requires_grad for _, p in named):
        return _ViTFeatures.apply(eng, x, [n for n, _ in named], *[p for _, p in named])
    return eng.features(x)


# --------------------------------------------------------------------------------------
# reference src/models/classifier.py
# --------------------------------------------------------------------------------------
class ClassificationHead(nn.Module):
    """Simple classification head."""

    def __init__(self, input_dim: int, output_dim: int) -> None:
        super().__init__()
        self.classification = nn.Linear(input_dim, output_dim)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.classification(x)


class ViTClassifier(nn.Module):
    """Classifier built on top of a pretrained ViT encoder (reference ``src/models/classifier.py:25-57``)."""

    def __init__(self, pretrained_encoder: nn.Module, num_classes: int = 10, head_cfg: Optional[Dict[str, Any]] = None):
        super().__init__()
        self.encoder = pretrained_encoder
        head_cfg = head_cfg or {}
        embed_dim = head_cfg.get("embed_dim", pretrained_encoder.embed_dim)
        self.pool_type = head_cfg.get("pool", "cls")  # or "mean"
        self.head = ClassificationHead(input_dim=embed_dim, output_dim=num_classes)

    def forward(self, x: torch.Tensor):
        feats = self.encoder.forward_features(x)
        if isinstance(feats, (tuple, list)):
            feats = feats[0]
        pooled = feats[:, 0] if self.pool_type == "cls" else feats.mean(dim=1)
        return self.head(pooled)


# --------------------------------------------------------------------------------------
# reference src/training/classifier.py
# --------------------------------------------------------------------------------------
class ViTClassifierTrainModule(_Base):
    """Supervised training for ViTClassifier: fine-tuning or linear probing (reference ``src/training/classifier.py:16``)."""

    def __init__(self, pretrained_encoder: Optional[nn.Module] = None, model_cfg: Optional[Dict[str, Any]] = None,
                 training_cfg: Optional[Dict[str, Any]] = None, num_classes: int = 10):
        super().__init__()
        self.save_hyperparameters(ignore=["pretrained_encoder"])
        self.model_cfg = model_cfg or {}
        self.training_cfg = training_cfg or {}
        tc = self.training_cfg
        self.learning_rate = float(tc.get("learning_rate", 3e-4))
        self.weight_decay = float(tc.get("weight_decay", 0.05))
        self.warmup_epochs = int(tc.get("warmup_epochs", 5))
        self.total_epochs = int(tc.get("total_epochs", 100))
        self.freeze_encoder_flag = tc.get("freeze_encoder", True)
        self.num_classes = num_classes
        enc_cfg = self.model_cfg.get("encoder", {})
        encoder = pretrained_encoder if pretrained_encoder is not None else VisionTransformer(
            img_size=self.model_cfg["general"]["image_size"], patch_size=self.model_cfg["general"]["patch_size"],
            in_chans=self.model_cfg["general"]["in_chans"], embed_dim=enc_cfg.get("embed_dim", 384),
            depth=enc_cfg.get("depth", 12), num_heads=enc_cfg.get("num_heads", 6), num_classes=0)
        self.model = ViTClassifier(pretrained_encoder=encoder, num_classes=num_classes,
                                   head_cfg=self.model_cfg.get("head", {}))
        if self.freeze_encoder_flag:
            self.freeze_encoder()
        else:
            self.unfreeze_encoder()

    def forward(self, x: torch.Tensor):
        return self.model(x)

    def _step(self, batch, prefix: str):
        imgs, labels = batch
        logits = self(imgs)
        loss = F.cross_entropy(logits, labels)
        acc = (logits.argmax(dim=1) == labels).float().mean()
        self.log(f"{prefix}_loss", loss, prog_bar=True)
        self.log(f"{prefix}_acc", acc, prog_bar=True)
        return loss

    def training_step(self, batch, batch_idx):
        return self._step(batch, "train")

    def validation_step(self, batch, batch_idx):
        return self._step(batch, "val")

    def test_step(self, batch, batch_idx):
        return self._step(batch, "test")

    def configure_optimizers(self):
        optimizer = AdamW(filter(lambda p: p.requires_grad, self.model.parameters()), lr=self.learning_rate,
                          weight_decay=self.weight_decay)
        scheduler = LambdaLR(optimizer, lambda e: lr_factor(e, self.warmup_epochs, self.total_epochs))
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": "epoch", "monitor": "val_loss"}}

    def freeze_encoder(self):
        for name, param in self.model.named_parameters():
            if "head" not in name:
                param.requires_grad = False

    def unfreeze_encoder(self):
        for param in self.model.parameters():
            param.requires_grad = True

    def unfreeze_last_layers(self, n_layers: int):
        """Unfreeze only the last ``n_layers`` blocks of the encoder (+ final norm + head); earlier layers stay frozen."""
        encoder = self.model.encoder
        blocks = encoder.blocks
        total = len(blocks)
        if n_layers < 0 or n_layers > total:
            raise ValueError(f"n_layers must be between 0 and {total}, got {n_layers}")
        for param in encoder.parameters():
            param.requires_grad = False
        for block in blocks[total - n_layers:]:
            for param in block.parameters():
                param.requires_grad = True
        if hasattr(encoder, "norm"):
            for param in encoder.norm.parameters():
                param.requires_grad = True
        for param in self.model.head.parameters():
            param.requires_grad = True
