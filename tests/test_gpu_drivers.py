"""-m gpu: the CLI drivers end to end on synthetic STL-10-shaped data: pretrain (2 epochs) -> outputs layout and
Lightning-shaped checkpoints -> resume -> classifier probe from the saved encoder."""
import os

import numpy as np
import pytest
import torch
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_pretrain_resume_and_finetune_drivers(tmp_path, monkeypatch):
    from scripts.training import pretrain_mae, train_mae
    monkeypatch.setenv("MAE_SYNTHETIC_DATA", "1")
    monkeypatch.setenv("MAE_SYNTHETIC_SIZE", "160")
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "mae.yaml")))
    cfg["pretrain"].update(batch_size=32, num_workers=0, total_epochs=2, warmup_epochs=1, val_split=0.2)
    cfg["train"].update(batch_size=32, num_workers=0, total_epochs=1, samples_per_class=8, freeze_encoder=True)
    cfg["logging"]["output_dir_base"] = str(tmp_path / "outputs")
    cfg_path = tmp_path / "cfg.yaml"
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    tr = pretrain_mae.main(["--config", str(cfg_path), "--output_dir_suffix", "t"])
    out = tmp_path / "outputs" / "pretrain" / "t"
    assert (out / "config.yaml").exists() and (out / "vit-mae.pt").exists()
    assert (out / "checkpoints" / "best.ckpt").exists() and (out / "checkpoints" / "last.ckpt").exists()
    assert len(tr.history) == 2 and all(torch.isfinite(torch.tensor(h["val_loss"])) for h in tr.history)
    assert tr.history[1]["train_loss"] < tr.history[0]["train_loss"] * 1.2
    ck = torch.load(out / "checkpoints" / "last.ckpt", weights_only=False)
    assert ck["epoch"] == 1 and all(k.startswith("model.") for k in ck["state_dict"]) and len(ck["state_dict"]) == 87
    assert any(v["monitor"] == "val_loss" and v["best_model_score"] is not None for v in ck["callbacks"].values())
    sd = torch.load(out / "vit-mae.pt")
    assert len(sd) == 87 and "encoder.vit.blocks.3.mlp.fc2.weight" in sd
    # resume: one more epoch, optimizer state restored
    cfg["pretrain"]["total_epochs"] = 3
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    tr2 = pretrain_mae.main(["--config", str(cfg_path), "--output_dir_suffix", "t", "--resume_from",
                             str(out / "checkpoints" / "last.ckpt")])
    assert [h["epoch"] for h in tr2.history] == [2] and tr2._fused.engine.step_count == tr.global_step + tr2.global_step - tr.global_step
    assert tr2._fused.engine.step_count > tr._fused.engine.step_count
    sd = torch.load(out / "vit-mae.pt")  # rewritten by the resumed run
    # downstream probe from the saved encoder weights (prefix "encoder.")
    tr3 = train_mae.main(["--config", str(cfg_path), "--encoder_ckpt", str(out / "vit-mae.pt"), "--output_dir_suffix", "p"])
    out2 = tmp_path / "outputs" / "train" / "p"
    assert (out2 / "vit-mae.pt").exists() and (out2 / "checkpoints" / "last.ckpt").exists()
    assert 0.0 <= tr3.history[0]["val_acc"] <= 1.0
    enc_before = sd["encoder.vit.blocks.0.attn.qkv.weight"]
    enc_after = torch.load(out2 / "vit-mae.pt")["encoder.blocks.0.attn.qkv.weight"]
    assert torch.equal(enc_before.cpu(), enc_after.cpu())  # frozen encoder untouched by the probe


def test_pretrain_driver_with_gpu_augment(tmp_path, monkeypatch):
    """``pretrain.gpu_augment: true``: uint8 batches from the loader, crop / flip / normalize on the GPU."""
    from scripts.training import pretrain_mae
    monkeypatch.setenv("MAE_SYNTHETIC_DATA", "1")
    monkeypatch.setenv("MAE_SYNTHETIC_SIZE", "160")
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "mae.yaml")))
    cfg["pretrain"].update(batch_size=32, num_workers=0, total_epochs=2, warmup_epochs=1, val_split=0.2, gpu_augment=True)
    cfg["logging"]["output_dir_base"] = str(tmp_path / "outputs")
    cfg_path = tmp_path / "cfg.yaml"
    yaml.safe_dump(cfg, open(cfg_path, "w"))
    tr = pretrain_mae.main(["--config", str(cfg_path), "--output_dir_suffix", "g"])
    assert getattr(tr, "_aug_train", None) is not None and getattr(tr, "_aug_eval", None) is not None
    assert len(tr.history) == 2 and all(np.isfinite(h["val_loss"]) and np.isfinite(h["train_loss"]) for h in tr.history)


def test_lightning_surface_fit_checkpoint_and_test(tmp_path, monkeypatch):
    """The calls the reference's scripts make on pytorch_lightning (``pl.Trainer(...).fit / .test``, ``ModelCheckpoint``,
    ``TensorBoardLogger``, ``load_from_checkpoint``; ``scripts/training/pretrain_mae.py:83-128``, ``scripts/utils.py:186-199``)
    executed for real on the GPU through the stand-in package (or the real one when installed)."""
    import pytorch_lightning as pl
    from pytorch_lightning.callbacks import LearningRateMonitor, ModelCheckpoint
    from pytorch_lightning.loggers import TensorBoardLogger

    from src.data import get_pretrain_dataloaders, get_test_dataloader, get_train_dataloaders
    from src.training.classifier import ViTClassifierTrainModule
    from src.training.mae import MAEPretrainModule
    monkeypatch.setenv("MAE_SYNTHETIC_DATA", "1")
    monkeypatch.setenv("MAE_SYNTHETIC_SIZE", "128")
    pl.seed_everything(73, workers=True)
    cfg = yaml.safe_load(open(os.path.join(ROOT, "configs", "mae.yaml")))
    cfg["pretrain"].update(batch_size=32, num_workers=0, total_epochs=2, warmup_epochs=1, val_split=0.25)
    cfg["train"].update(batch_size=32, num_workers=0, total_epochs=1, samples_per_class=6, freeze_encoder=False)
    cfg["test"].update(batch_size=32, num_workers=0)
    tl, vl = get_pretrain_dataloaders(cfg)
    module = MAEPretrainModule(model_cfg=cfg["model"], training_cfg=cfg["pretrain"])
    ckpt_dir = tmp_path / "ck"
    logger = TensorBoardLogger(save_dir=str(tmp_path / "logs"), name="tb")
    best = ModelCheckpoint(dirpath=ckpt_dir, filename="best", save_top_k=1, monitor="val_loss", mode="min", save_last=False, verbose=True)
    last = ModelCheckpoint(dirpath=ckpt_dir, filename="last", save_top_k=1, every_n_epochs=1, verbose=True)
    trainer = pl.Trainer(accelerator="auto", devices=1, max_epochs=2, logger=logger, callbacks=[best, last, LearningRateMonitor("epoch")],
                         log_every_n_steps=2, precision="bf16-mixed", gradient_clip_val=1.0, gradient_clip_algorithm="norm")
    trainer.fit(module, tl, vl, ckpt_path=None)
    assert os.path.exists(best.best_model_path) and (ckpt_dir / "last.ckpt").exists() and os.path.isdir(logger.log_dir)
    ck = torch.load(ckpt_dir / "last.ckpt", weights_only=False)
    assert ck["pytorch-lightning_version"] == "2.5.6" and ck["epoch"] == 1 and trainer.current_epoch == 1
    # classifier: fit one epoch, reload with load_from_checkpoint(strict=False), trainer.test
    tl, vl = get_train_dataloaders(cfg)
    cls = ViTClassifierTrainModule(pretrained_encoder=module.model.encoder.vit, model_cfg=cfg["model"], training_cfg=cfg["train"])
    cbest = ModelCheckpoint(dirpath=tmp_path / "ck2", filename="best", monitor="val_acc", mode="max")
    pl.Trainer(max_epochs=1, callbacks=[cbest], precision="bf16-mixed", gradient_clip_val=1.0).fit(cls, tl, vl)
    again = ViTClassifierTrainModule.load_from_checkpoint(cbest.best_model_path, map_location="cpu", strict=False)
    assert torch.equal(again.model.head.classification.weight.cpu(), cls.model.head.classification.weight.detach().cpu())
    res = pl.Trainer(accelerator="auto", devices=1, precision="bf16-mixed").test(again, get_test_dataloader(cfg), ckpt_path=None)
    assert isinstance(res, list) and 0.0 <= res[0]["test_acc"] <= 1.0 and np.isfinite(res[0]["test_loss"])
