"""Self-supervised MAE pre-training driver: same CLI, YAML sections and output layout as the reference's
``scripts/training/pretrain_mae.py`` (``outputs/pretrain/<suffix>/{config.yaml, checkpoints/{best,last,epoch-XXX}.ckpt,
vit-mae.pt}``), running the B200-native fused step.   python -m scripts.training.pretrain_mae --config configs/mae.yaml"""
from __future__ import annotations

import argparse
from pathlib import Path

import torch
import yaml

from src.data import get_pretrain_dataloaders
from src.training.mae import MAEPretrainModule
from ssrl_vit_mae_jepa_b200.trainer import LiteTrainer, ModelCheckpoint

from ..utils import setup_reproducibility, shut_down_warnings


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Self-supervised MAE pretraining")
    p.add_argument("--config", type=str, default="configs/mae.yaml")
    p.add_argument("--resume_from", type=str, default=None, help="Path to checkpoint to resume from")
    p.add_argument("--output_dir_suffix", type=str, default="mae_pretrain", help="Suffix for the output directory")
    p.add_argument("--limit_batches", type=int, default=None, help="(extra) cap batches per epoch, for smoke runs")
    return p.parse_args(argv)


def main(argv=None):
    shut_down_warnings()
    setup_reproducibility(seed=73)
    args = parse_args(argv)
    with open(args.config) as f:
        cfg = yaml.safe_load(f)
    pre_cfg, model_cfg, log_cfg = cfg["pretrain"], cfg["model"], cfg["logging"]
    output_dir = Path(log_cfg["output_dir_base"]) / "pretrain" / args.output_dir_suffix
    ckpt_dir = output_dir / "checkpoints"
    ckpt_dir.mkdir(parents=True, exist_ok=True)
    with open(output_dir / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    train_loader, val_loader = get_pretrain_dataloaders(cfg)
    module = MAEPretrainModule(model_cfg=model_cfg, training_cfg=pre_cfg)
    ckpt_best = ModelCheckpoint(ckpt_dir, "best", monitor="val_loss", mode="min")
    ckpt_last = ModelCheckpoint(ckpt_dir, "last", every_n_epochs=1)
    ckpt_periodic = ModelCheckpoint(ckpt_dir, "epoch-{epoch:03d}", every_n_epochs=25, save_weights_only=True, save_top_k=-1)
    trainer = LiteTrainer(max_epochs=pre_cfg["total_epochs"], callbacks=[ckpt_best, ckpt_last, ckpt_periodic],
                          gradient_clip_val=1.0, limit_batches=args.limit_batches)
    trainer.fit(module, train_loader, val_loader, ckpt_path=args.resume_from)
    model_path = output_dir / log_cfg["model_path"]
    torch.save(module.model.state_dict(), model_path)
    print(f"pretraining complete; weights: {model_path}; best checkpoint: {ckpt_best.best_model_path}")
    return trainer


if __name__ == "__main__":
    main()
