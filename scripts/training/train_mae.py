"""Downstream classifier training driver (frozen-encoder probe / unfreeze-last-N / full fine-tune): same CLI and YAML
sections as the reference's ``scripts/training/train_mae.py``.   python -m scripts.training.train_mae --encoder_ckpt ..."""
from __future__ import annotations

import argparse
from pathlib import Path

import torch
import yaml

from src.data import get_train_dataloaders
from src.models.mae import MaskedAutoencoder
from src.training.classifier import ViTClassifierTrainModule
from ssrl_vit_mae_jepa_b200.trainer import LiteTrainer, ModelCheckpoint

from ..utils import load_encoder_weights, setup_reproducibility, shut_down_warnings


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Fine-tune or train MAE encoder on classification task")
    p.add_argument("--config", type=str, default="configs/mae.yaml")
    p.add_argument("--encoder_ckpt", type=str, default=None, help="Path to pretrained MAE encoder weights (.pt or .ckpt)")
    p.add_argument("--classifier_ckpt", type=str, default=None, help="Path to full classifier checkpoint")
    p.add_argument("--output_dir_suffix", type=str, default="mae_finetune", help="Suffix for the output directory")
    p.add_argument("--limit_batches", type=int, default=None, help="(extra) cap batches per epoch, for smoke runs")
    return p.parse_args(argv)


def main(argv=None):
    shut_down_warnings()
    setup_reproducibility(seed=73)
    args = parse_args(argv)
    with open(args.config) as f:
        cfg = yaml.safe_load(f)
    model_cfg, train_cfg, log_cfg = cfg["model"], cfg["train"], cfg["logging"]
    output_dir = Path(log_cfg["output_dir_base"]) / "train" / args.output_dir_suffix
    ckpt_dir = output_dir / "checkpoints"
    ckpt_dir.mkdir(parents=True, exist_ok=True)
    with open(output_dir / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    train_loader, val_loader = get_train_dataloaders(cfg)
    if args.classifier_ckpt is not None:
        # weight initialisation from a full classifier checkpoint, like the reference's load_from_checkpoint(strict=False)
        # (scripts/training/train_mae.py:85-91): takes precedence over --encoder_ckpt, training starts at epoch 0
        module = ViTClassifierTrainModule.load_from_checkpoint(args.classifier_ckpt, map_location="cpu", strict=False)
    elif args.encoder_ckpt is not None:
        mae = MaskedAutoencoder(model_cfg["general"], model_cfg["encoder"], model_cfg["decoder"])
        load_encoder_weights(mae.encoder, args.encoder_ckpt)
        module = ViTClassifierTrainModule(pretrained_encoder=mae.encoder.vit, model_cfg=model_cfg, training_cfg=train_cfg)
    else:  # randomly initialised ViT baseline
        module = ViTClassifierTrainModule(model_cfg=model_cfg, training_cfg=train_cfg)
    if train_cfg.get("unfreeze_last_layers") is not None:
        module.unfreeze_last_layers(int(train_cfg["unfreeze_last_layers"]))
    elif train_cfg.get("freeze_encoder", True):
        module.freeze_encoder()
    else:
        module.unfreeze_encoder()
    ckpt_best = ModelCheckpoint(ckpt_dir, "best", monitor="val_acc", mode="max")
    ckpt_last = ModelCheckpoint(ckpt_dir, "last", every_n_epochs=1)
    trainer = LiteTrainer(max_epochs=train_cfg["total_epochs"], callbacks=[ckpt_best, ckpt_last], gradient_clip_val=1.0,
                          limit_batches=args.limit_batches)
    trainer.fit(module, train_loader, val_loader)
    model_path = output_dir / log_cfg["model_path"]
    torch.save(module.model.state_dict(), model_path)
    print(f"training complete; weights: {model_path}; best checkpoint: {ckpt_best.best_model_path}")
    return trainer


if __name__ == "__main__":
    main()
