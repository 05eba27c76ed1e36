"""Helpers shared by the CLI drivers (reference ``scripts/utils.py:13-38``)."""
import random
import warnings

import numpy as np
import torch


def setup_reproducibility(seed: int) -> None:
    """Fix the random seeds (the reference calls ``pl.seed_everything(seed)``)."""
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.benchmark = False


def shut_down_warnings() -> None:
    warnings.filterwarnings("ignore", "Precision bf16-mixed is not supported by the model summary")
    warnings.filterwarnings("ignore", "Please use the new API settings to control TF32 behavior")


def load_encoder_weights(encoder: torch.nn.Module, path: str) -> str:
    """Load ViT encoder weights from a ``vit-mae.pt`` state_dict or a ``.ckpt``: detect where the encoder lives
    (``model.encoder.`` LightningModule, ``encoder.`` plain MAE state_dict, ``module.encoder.`` DDP), strip the
    prefix and load non-strictly (reference ``scripts/training/train_mae.py:100-135``).  ``encoder`` is the
    ``MaskedAutoencoder.encoder`` wrapper (keys ``mask_token``, ``vit.*``).  Returns the detected prefix."""
    ckpt = torch.load(path, map_location="cpu", weights_only=False)
    sd = ckpt.get("state_dict", ckpt)
    prefixes = ["model.encoder.", "encoder.", "module.encoder."]
    prefix = next((p for p in prefixes if any(k.startswith(p) for k in sd)), None)
    if prefix is None:
        raise RuntimeError("no encoder weights in checkpoint; expected keys starting with one of: " + ", ".join(prefixes))
    enc = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}
    missing, unexpected = encoder.load_state_dict(enc, strict=False)
    print(f"[ckpt] encoder prefix '{prefix}': {len(enc)} tensors, missing {len(missing)}, unexpected {len(unexpected)}")
    return prefix
