"""Generate the committed golden fixtures for the MAE hot path and pin the oracle.

Run from the repo root:  ``python tests/golden/make_golden.py``

The reference (``/root/reference/src/models/mae.py``) cannot be imported in this image (timm,
lightly and pytorch_lightning are not installed and there is no network), so the vectors are
produced by ``oracle/mae_oracle.py`` and are only written after the oracle has been
cross-checked against an INDEPENDENT implementation that is installed here: HuggingFace
``transformers.models.vit_mae.ViTMAEForPreTraining`` with the oracle's weights mapped across and
the same masking noise.  Agreement required: loss 1e-6 relative, predictions 2e-5 absolute.

Outputs (small, committed): ``tests/golden/mae_b4_seed73.npz``.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import mae_oracle as O  # noqa: E402


def hf_twin(model: O.MaskedAutoencoder):
    """Build a HuggingFace ViTMAE with the oracle's weights.  HF keeps 36 patches + CLS at ratio 0.75
    (37 tokens) whereas lightly keeps 36 tokens INCLUDING CLS, so the HF ratio is chosen
    to give int(144*(1-r)) == num_keep-1 patches."""
    from transformers import ViTMAEConfig, ViTMAEForPreTraining

    L = model.sequence_length
    keep_patches = O.num_keep_tokens(L, model.mask_ratio) - 1
    r = 1.0 - (keep_patches + 0.5) / (L - 1)
    assert int((L - 1) * (1 - r)) == keep_patches
    enc = model.encoder.vit
    cfg = ViTMAEConfig(
        hidden_size=enc.embed_dim, num_hidden_layers=len(enc.blocks), num_attention_heads=enc.blocks[0].attn.num_heads,
        intermediate_size=enc.embed_dim * 4, image_size=model.image_size, patch_size=model.patch_size,
        num_channels=model.in_chans, decoder_hidden_size=model.decoder.decoder_embed.out_features,
        decoder_num_hidden_layers=len(model.decoder.decoder_blocks),
        decoder_num_attention_heads=model.decoder.decoder_blocks[0].attn.num_heads,
        decoder_intermediate_size=model.decoder.decoder_embed.out_features * 4, mask_ratio=r, layer_norm_eps=1e-6,
        norm_pix_loss=False, hidden_act="gelu", qkv_bias=True, attn_implementation="eager")
    hf = ViTMAEForPreTraining(cfg).eval()
    sd = model.state_dict()
    out = {}
    out["vit.embeddings.cls_token"] = sd["encoder.vit.cls_token"]
    out["vit.embeddings.position_embeddings"] = sd["encoder.vit.pos_embed"]
    out["vit.embeddings.patch_embeddings.projection.weight"] = sd["encoder.vit.patch_embed.proj.weight"]
    out["vit.embeddings.patch_embeddings.projection.bias"] = sd["encoder.vit.patch_embed.proj.bias"]

    def block(src: str, dst: str):
        D = sd[src + "attn.proj.weight"].shape[0]
        qkv_w, qkv_b = sd[src + "attn.qkv.weight"], sd[src + "attn.qkv.bias"]
        for i, n in enumerate(["query", "key", "value"]):
            out[dst + f"attention.attention.{n}.weight"] = qkv_w[i * D:(i + 1) * D]
            out[dst + f"attention.attention.{n}.bias"] = qkv_b[i * D:(i + 1) * D]
        for a, b in [("attn.proj", "attention.output.dense"), ("norm1", "layernorm_before"),
                     ("norm2", "layernorm_after"), ("mlp.fc1", "intermediate.dense"), ("mlp.fc2", "output.dense")]:
            out[dst + b + ".weight"] = sd[src + a + ".weight"]
            out[dst + b + ".bias"] = sd[src + a + ".bias"]

    for i in range(len(enc.blocks)):
        block(f"encoder.vit.blocks.{i}.", f"vit.encoder.layer.{i}.")
    out["vit.layernorm.weight"] = sd["encoder.vit.norm.weight"]
    out["vit.layernorm.bias"] = sd["encoder.vit.norm.bias"]
    out["decoder.mask_token"] = sd["decoder.mask_token"]
    out["decoder.decoder_pos_embed"] = sd["decoder.decoder_pos_embed"]
    for n in ["decoder_embed", "decoder_norm", "decoder_pred"]:
        out[f"decoder.{n}.weight"] = sd[f"decoder.{n}.weight"]
        out[f"decoder.{n}.bias"] = sd[f"decoder.{n}.bias"]
    for i in range(len(model.decoder.decoder_blocks)):
        block(f"decoder.decoder_blocks.{i}.", f"decoder.decoder_layers.{i}.")
    missing, unexpected = hf.load_state_dict(out, strict=True), None
    return hf


def cross_check(model: O.MaskedAutoencoder, images: torch.Tensor, noise: torch.Tensor):
    hf = hf_twin(model)
    with torch.no_grad():
        pred, target = model(images, noise=noise)
        loss = torch.nn.functional.mse_loss(pred, target)
        idx_keep, idx_mask = O.random_token_mask_from_noise(noise, model.mask_ratio)
        o = hf(pixel_values=images, noise=noise[:, 1:].contiguous())
    hf_pred = O.get_at_index(o.logits, idx_mask - 1)
    # HF's binary mask (1 = masked) must equal the complement of idx_keep
    mask = torch.ones(images.shape[0], model.sequence_length - 1)
    mask.scatter_(1, idx_keep[:, 1:] - 1, 0.0)
    assert torch.equal(mask, o.mask), "mask mismatch vs HF ViTMAE"
    rel = abs(o.loss.item() - loss.item()) / abs(loss.item())
    err = (hf_pred - pred).abs().max().item()
    assert rel < 1e-6, f"loss mismatch vs HF: {o.loss.item()} vs {loss.item()}"
    assert err < 2e-5, f"pred mismatch vs HF: {err}"
    # patchify and sin-cos tables against HF's own helpers
    assert torch.equal(hf.patchify(images), O.patchify(images, model.patch_size))
    from transformers.models.vit_mae.modeling_vit_mae import get_2d_sincos_pos_embed
    for D in (model.encoder.vit.embed_dim, model.decoder.decoder_embed.out_features):
        ref = torch.from_numpy(get_2d_sincos_pos_embed(D, 12, add_cls_token=True)).float()
        assert (ref - O.sincos_2d_pos_embed(D, 12)[0]).abs().max() < 1e-6
    return rel, err


def main():
    torch.set_float32_matmul_precision("highest")
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    model = O.build_mae(seed=73)
    g = torch.Generator().manual_seed(173)
    B = 4
    images = torch.randn(B, 3, 96, 96, generator=g)
    noise = torch.rand(B, model.sequence_length, generator=g)
    rel, err = cross_check(model, images, noise)
    print(f"HF ViTMAE cross-check OK: loss rel diff {rel:.2e}, pred max abs diff {err:.2e}")

    loss, pred, target, grads = O.loss_and_grads(model, images, noise)
    idx_keep, idx_mask = O.random_token_mask_from_noise(noise, model.mask_ratio)
    out = {
        "images": images.numpy(), "noise": noise.numpy(), "idx_keep": idx_keep.numpy(), "idx_mask": idx_mask.numpy(),
        "loss": np.float32(loss.item()), "pred": pred.numpy().astype(np.float32),
        "target_checksum": np.float64(target.double().abs().sum().item()),
        "weight_checksum": np.float64(sum(v.double().abs().sum().item() for v in model.state_dict().values())),
    }
    for name in ["decoder.decoder_pred.weight", "decoder.mask_token", "encoder.vit.cls_token",
                 "encoder.vit.blocks.0.attn.qkv.weight", "encoder.vit.patch_embed.proj.weight",
                 "decoder.decoder_blocks.1.mlp.fc2.bias", "encoder.vit.norm.weight"]:
        out["grad:" + name] = grads[name].numpy()
    out["grad_norms"] = np.array([grads[n].double().norm().item() for n in sorted(grads)], dtype=np.float64)
    out["grad_names"] = np.array(sorted(grads))
    # a tie case for the stable-sort rule
    tie = noise[:2].clone()
    tie[0, 5] = tie[0, 90]
    tie[0, 17] = tie[0, 3]
    tie[1, 1:] = 0.25
    k, m = O.random_token_mask_from_noise(tie, 0.75)
    out["tie_noise"], out["tie_idx_keep"], out["tie_idx_mask"] = tie.numpy(), k.numpy(), m.numpy()
    path = os.path.join(ROOT, "tests", "golden", "mae_b4_seed73.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; loss =", loss.item())


if __name__ == "__main__":
    main()
