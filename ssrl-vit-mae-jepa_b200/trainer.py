"""A small Lightning-style fit loop for the two training modules, with the reference's checkpoint policy
(``scripts/training/pretrain_mae.py:85-128``): ``best`` (monitored metric), ``last`` (every epoch), periodic
weights-only ``epoch-XXX`` files, ``--resume_from`` via ``fit(ckpt_path=...)``, and Lightning-shaped ``.ckpt``
dictionaries (``state_dict`` with the ``model.`` prefix the reference's loaders strip, ``epoch``, ``global_step``,
``callbacks[...]["best_model_score"]``).

``MAEPretrainModule`` trains through ``FusedTrainer`` (fused clip+AdamW on the flat buffers, no autograd graph);
any other module (the classifier) trains through autograd with the optimizer/scheduler it configures, gradient
clipping 1.0 by norm, exactly the reference's Trainer settings.
"""
from __future__ import annotations

import os
from typing import Any, Dict, List, Optional

import torch

from .training import DevicePrefetcher, FusedTrainer, MAEPretrainModule


class ModelCheckpoint:
    def __init__(self, dirpath, filename: str, monitor: Optional[str] = None, mode: str = "min", every_n_epochs: int = 1,
                 save_weights_only: bool = False, save_top_k: int = 1, **_):
        self.dirpath, self.filename, self.monitor, self.mode = str(dirpath), filename, monitor, mode
        self.every_n_epochs, self.save_weights_only, self.save_top_k = every_n_epochs, save_weights_only, save_top_k
        self.best_model_score: Optional[torch.Tensor] = None
        self.best_model_path = ""

    def state_key(self) -> str:
        return f"ModelCheckpoint{{'monitor': {self.monitor!r}, 'mode': {self.mode!r}, 'filename': {self.filename!r}}}"

    def maybe_save(self, trainer: "LiteTrainer", epoch: int, metrics: Dict[str, float]) -> None:
        if (epoch + 1) % max(1, self.every_n_epochs):
            return
        if self.monitor is not None:
            if self.monitor not in metrics:
                return
            v = metrics[self.monitor]
            better = self.best_model_score is None or (v < self.best_model_score.item() if self.mode == "min"
                                                       else v > self.best_model_score.item())
            if not better:
                return
            self.best_model_score = torch.tensor(v)
        path = os.path.join(self.dirpath, self.filename.format(epoch=epoch) + ".ckpt")
        os.makedirs(self.dirpath, exist_ok=True)
        self.best_model_path = path
        torch.save(trainer.checkpoint_dict(weights_only=self.save_weights_only), path)


class LiteTrainer:
    def __init__(self, max_epochs: int, callbacks: Optional[List[Any]] = None, gradient_clip_val: float = 1.0,
                 device: str = "cuda", log_every_n_steps: int = 2, limit_batches: Optional[int] = None, **_):
        self.max_epochs, self.clip, self.device = max_epochs, gradient_clip_val, torch.device(device)
        self.callbacks = [c for c in (callbacks or []) if isinstance(c, ModelCheckpoint)]
        self.current_epoch, self.global_step = 0, 0
        self.limit_batches = limit_batches
        self.history: List[Dict[str, float]] = []
        self.module = None
        self._fused: Optional[FusedTrainer] = None
        self._opt = self._sched = None

    # -------- checkpoints --------
    def checkpoint_dict(self, weights_only: bool = False) -> Dict[str, Any]:
        m = self.module
        ck: Dict[str, Any] = {
            "epoch": self.current_epoch, "global_step": self.global_step, "pytorch-lightning_version": "lite",
            "state_dict": {k: v.detach().cpu().clone() for k, v in m.state_dict().items()},
            "hyper_parameters": {k: v for k, v in getattr(m, "hparams", {}).items() if isinstance(v, (dict, int, float, str, bool, type(None)))},
            "callbacks": {c.state_key(): {"best_model_score": c.best_model_score, "best_model_path": c.best_model_path,
                                          "monitor": c.monitor} for c in self.callbacks},
        }
        if not weights_only:
            if self._fused is not None:
                e = self._fused.engine
                ck["optimizer_states"] = [{"fused_adamw": True, "step": e.step_count, "exp_avg": e.exp_avg.cpu(),
                                           "exp_avg_sq": e.exp_avg_sq.cpu()}]
            elif self._opt is not None:
                ck["optimizer_states"] = [self._opt.state_dict()]
                ck["lr_schedulers"] = [self._sched.state_dict()] if self._sched is not None else []
        return ck

    def _restore(self, path: str) -> None:
        ck = torch.load(path, map_location="cpu", weights_only=False)
        self.module.load_state_dict(ck["state_dict"], strict=True)
        self.current_epoch, self.global_step = int(ck.get("epoch", -1)) + 1, int(ck.get("global_step", 0))
        for c in self.callbacks:
            st = ck.get("callbacks", {}).get(c.state_key())
            if st:
                c.best_model_score, c.best_model_path = st["best_model_score"], st["best_model_path"]
        self._pending_opt_state = (ck.get("optimizer_states") or [None])[0]
        print(f"[trainer] resumed from {path}: starting at epoch {self.current_epoch}")

    # -------- loop --------
    def fit(self, module, train_loader, val_loader=None, ckpt_path: Optional[str] = None) -> None:
        self.module = module.to(self.device)
        self._pending_opt_state = None
        if ckpt_path:
            self._restore(ckpt_path)
        fused = isinstance(module, MAEPretrainModule)
        if fused:
            self._fused = FusedTrainer(module, max_norm=self.clip)
        else:
            oc = module.configure_optimizers()
            self._opt, self._sched = oc["optimizer"], oc["lr_scheduler"]["scheduler"]
            for _ in range(self.current_epoch):
                self._sched.step()
        st = self._pending_opt_state
        for epoch in range(self.current_epoch, self.max_epochs):
            self.current_epoch = module.current_epoch = epoch
            metrics: Dict[str, float] = {}
            if fused:
                self._fused.set_epoch(epoch)
            elif hasattr(module, "on_train_epoch_start"):
                module.on_train_epoch_start()
            losses = []
            batches = DevicePrefetcher(self._images_only(train_loader), self.device) if fused else train_loader
            for bi, batch in enumerate(batches):
                if self.limit_batches is not None and bi >= self.limit_batches:
                    break
                if fused:
                    if st is not None and st.get("fused_adamw"):  # engine buffers exist after the first forward
                        self._fused.engine._ensure_flat(self.device)
                        e = self._fused.engine
                        e.exp_avg.copy_(st["exp_avg"]); e.exp_avg_sq.copy_(st["exp_avg_sq"]); e.step_count = int(st["step"])
                        st = None
                    losses.append(self._fused.train_step(self._prep(batch, True)))
                else:
                    if st is not None and not st.get("fused_adamw"):
                        self._opt.load_state_dict(st)
                        st = None
                    imgs, labels = batch
                    self._opt.zero_grad(set_to_none=True)
                    loss = module.training_step((self._prep(imgs.to(self.device, non_blocking=True), True), labels.to(self.device)), bi)
                    loss.backward()
                    torch.nn.utils.clip_grad_norm_([p for p in module.parameters() if p.requires_grad], self.clip)
                    self._opt.step()
                    losses.append(loss.detach())
                self.global_step += 1
            if losses:
                metrics["train_loss"] = torch.stack([l.float() for l in losses]).mean().item()
            if not fused and self._sched is not None:
                self._sched.step()
            if val_loader is not None:
                metrics.update(self._validate(module, val_loader))
            for k in ("train_acc", "mask_ratio"):
                if k in getattr(module, "logged", {}):
                    v = module.logged[k]
                    metrics[k] = v.item() if torch.is_tensor(v) else float(v)
            metrics["epoch"] = epoch
            self.history.append(metrics)
            print("[trainer] " + "  ".join(f"{k}={v:.4f}" if isinstance(v, float) else f"{k}={v}" for k, v in metrics.items()))
            for c in self.callbacks:
                c.maybe_save(self, epoch, metrics)

    def _prep(self, images: torch.Tensor, train: bool) -> torch.Tensor:
        """uint8 batches (``gpu_augment`` loaders) get the reference transform on the GPU; float batches pass through."""
        if images.dtype != torch.uint8:
            return images
        from .data import GpuAugment

        key = "_aug_train" if train else "_aug_eval"
        if getattr(self, key, None) is None:
            setattr(self, key, GpuAugment(size=images.shape[-1], train=train))
        return getattr(self, key)(images)

    @staticmethod
    def _images_only(loader):
        for batch in loader:
            yield batch[0] if isinstance(batch, (tuple, list)) else batch

    @torch.no_grad()
    def _validate(self, module, loader) -> Dict[str, float]:
        tot, n, correct = 0.0, 0, 0.0
        for bi, batch in enumerate(loader):
            if self.limit_batches is not None and bi >= self.limit_batches:
                break
            imgs, labels = batch
            b = (self._prep(imgs.to(self.device, non_blocking=True), False),
                 labels.to(self.device) if torch.is_tensor(labels) else labels)
            loss = module.validation_step(b, bi)
            tot += loss.item() * imgs.shape[0]
            n += imgs.shape[0]
            if "val_acc" in getattr(module, "logged", {}):
                correct += module.logged["val_acc"].item() * imgs.shape[0]
        out = {"val_loss": tot / max(n, 1)}
        if "val_acc" in getattr(module, "logged", {}):
            out["val_acc"] = correct / max(n, 1)
        return out
