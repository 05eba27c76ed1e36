"""A small Lightning-style fit loop for the two training modules, with the reference's checkpoint policy
(``scripts/training/pretrain_mae.py:85-128``): ``best`` (monitored metric), ``last`` (every epoch), periodic
weights-only ``epoch-XXX`` files, ``--resume_from`` via ``fit(ckpt_path=...)``, and Lightning-shaped ``.ckpt``
dictionaries (``state_dict`` with the ``model.`` prefix the reference's loaders strip, ``epoch``, ``global_step``,
``callbacks[...]["best_model_score"]``).

``MAEPretrainModule`` trains through ``FusedTrainer`` (fused clip+AdamW on the flat buffers, no autograd graph);
any other module (the classifier) trains through autograd with the optimizer/scheduler it configures, gradient
clipping 1.0 by norm, exactly the reference's Trainer settings.

Data parallelism (BASELINE config #3; the reference itself is ``devices=1``, ``scripts/training/pretrain_mae.py:116-118``):
launched under ``torchrun`` (``WORLD_SIZE > 1``) the trainer joins the NCCL process group, one process per GPU, broadcasts
rank 0's parameters, re-wraps the loaders with a ``DistributedSampler`` (every rank keeps the YAML's per-process batch
size, as ``effective_lr`` assumes), averages gradients (bucketed all-reduce overlapped with backward in ``FusedTrainer``;
one flat all-reduce for the autograd path), averages the logged metrics over ranks and writes checkpoints on rank 0 only.
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
        self.best_model_path = path
        if trainer.is_global_zero:  # every rank tracks the score; one writes
            os.makedirs(self.dirpath, exist_ok=True)
            torch.save(trainer.checkpoint_dict(weights_only=self.save_weights_only), path)


class LiteTrainer:
    def __init__(self, max_epochs: int, callbacks: Optional[List[Any]] = None, gradient_clip_val: float = 1.0,
                 device: str = "cuda", log_every_n_steps: int = 2, limit_batches: Optional[int] = None, logger=None, **_):
        self.max_epochs, self.clip, self.device = max_epochs, gradient_clip_val, torch.device(device)
        self.callbacks = [c for c in (callbacks or []) if isinstance(c, ModelCheckpoint)]
        self.current_epoch, self.global_step = 0, 0
        self.limit_batches = limit_batches
        self.logger = logger
        self.history: List[Dict[str, float]] = []
        self.module = None
        self._fused: Optional[FusedTrainer] = None
        self._opt = self._sched = None
        self.world, self.rank = 1, 0

    # -------- data parallelism --------
    @property
    def is_global_zero(self) -> bool:
        return self.rank == 0

    def _dist_setup(self) -> None:
        import torch.distributed as dist

        world = int(os.environ.get("WORLD_SIZE", "1"))
        if dist.is_available() and dist.is_initialized():
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
        elif world > 1 and dist.is_available():
            if self.device.type == "cuda":
                local = int(os.environ.get("LOCAL_RANK", "0"))
                torch.cuda.set_device(local)
                self.device = torch.device("cuda", local)
                dist.init_process_group("nccl", device_id=self.device)
            else:
                dist.init_process_group("gloo")
            self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if self.world > 1 and self.device.type == "cuda" and self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())

    def _shard(self, loader, seed: int = 73):
        """The same loader over this rank's shard (``DistributedSampler``); single process: unchanged."""
        from torch.utils.data import DataLoader, DistributedSampler, RandomSampler

        if self.world == 1 or loader is None or not hasattr(loader, "dataset") or \
                isinstance(getattr(loader, "sampler", None), DistributedSampler):
            return loader
        sampler = DistributedSampler(loader.dataset, num_replicas=self.world, rank=self.rank,
                                     shuffle=isinstance(loader.sampler, RandomSampler), seed=seed, drop_last=False)
        return DataLoader(loader.dataset, batch_size=loader.batch_size, sampler=sampler, num_workers=loader.num_workers,
                          pin_memory=True, drop_last=loader.drop_last, collate_fn=loader.collate_fn,
                          persistent_workers=loader.num_workers > 0)

    def _sync_params(self, module) -> None:
        if self.world > 1:
            import torch.distributed as dist

            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.data, 0)

    def _mean_over_ranks(self, *vals: float) -> List[float]:
        if self.world == 1:
            return list(vals)
        import torch.distributed as dist

        t = torch.tensor(vals, dtype=torch.float64, device=self.device)
        dist.all_reduce(t)
        return (t / self.world).tolist()

    def _allreduce_grads(self, params) -> None:
        """autograd path: one flat all-reduce(avg) of every gradient (the classifier is a few hundred KB to 4 MB)."""
        import torch.distributed as dist

        gs = [p.grad for p in params if p.grad is not None]
        if not gs:
            return
        flat = torch.cat([g.reshape(-1) for g in gs])
        dist.all_reduce(flat)
        flat.div_(self.world)
        o = 0
        for g in gs:
            g.copy_(flat[o:o + g.numel()].view_as(g))
            o += g.numel()

    def _log(self, metrics: Dict[str, float], step: int) -> None:
        if self.logger is not None and self.is_global_zero and hasattr(self.logger, "log_metrics"):
            self.logger.log_metrics({k: v for k, v in metrics.items() if isinstance(v, (int, float))}, step)

    # -------- checkpoints --------
    def checkpoint_dict(self, weights_only: bool = False) -> Dict[str, Any]:
        m = self.module
        ck: Dict[str, Any] = {
            "epoch": self.current_epoch, "global_step": self.global_step, "pytorch-lightning_version": "2.5.6",
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
        self._dist_setup()
        self.module = module.to(self.device)
        self._pending_opt_state = None
        if ckpt_path:
            self._restore(ckpt_path)
        self._sync_params(module)
        train_loader, val_loader = self._shard(train_loader), self._shard(val_loader)
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
            if hasattr(getattr(train_loader, "sampler", None), "set_epoch"):
                train_loader.sampler.set_epoch(epoch)
            metrics: Dict[str, float] = {}
            if fused:
                self._fused.set_epoch(epoch)
            elif hasattr(module, "on_train_epoch_start"):
                module.on_train_epoch_start()
            losses = []
            # uint8 batches (gpu_augment loaders) are cropped / flipped / normalised on the prefetch stream, behind their copy
            batches = DevicePrefetcher(self._images_only(train_loader), self.device, transform=self._augmenter(True)) \
                if fused else train_loader
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
                    trainable = [p for p in module.parameters() if p.requires_grad]
                    if self.world > 1:
                        self._allreduce_grads(trainable)
                    if self.clip:
                        torch.nn.utils.clip_grad_norm_(trainable, self.clip)
                    self._opt.step()
                    losses.append(loss.detach())
                self.global_step += 1
            if losses:
                metrics["train_loss"] = self._mean_over_ranks(torch.stack([l.float() for l in losses]).mean().item())[0]
            metrics["lr"] = self._fused.current_lr() if fused else self._opt.param_groups[0]["lr"]
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
            self._log(metrics, self.global_step)
            if self.is_global_zero:
                print("[trainer] " + "  ".join(f"{k}={v:.4f}" if isinstance(v, float) else f"{k}={v}" for k, v in metrics.items()))
            for c in self.callbacks:
                c.maybe_save(self, epoch, metrics)
        if self.logger is not None and hasattr(self.logger, "finalize"):
            self.logger.finalize()

    @torch.no_grad()
    def test(self, module, loader, ckpt_path: Optional[str] = None, prefix: str = "test") -> List[Dict[str, float]]:
        """``trainer.test`` (reference ``scripts/utils.py:196-199``): ``test_step`` over the loader, metrics averaged over
        samples (and ranks).  Returns ``[{"test_loss": ..., "test_acc": ...}]`` like Lightning."""
        self._dist_setup()
        self.module = module.to(self.device)
        if ckpt_path:
            ck = torch.load(ckpt_path, map_location="cpu", weights_only=False)
            module.load_state_dict(ck.get("state_dict", ck), strict=False)
        loader = self._shard(loader)
        step = getattr(module, f"{prefix}_step", None) or module.validation_step
        sums: Dict[str, float] = {}
        n = 0
        for bi, batch in enumerate(loader):
            if self.limit_batches is not None and bi >= self.limit_batches:
                break
            imgs, labels = batch
            module.logged = {}
            step((self._prep(imgs.to(self.device, non_blocking=True), False),
                  labels.to(self.device) if torch.is_tensor(labels) else labels), bi)
            for k, v in module.logged.items():
                sums[k] = sums.get(k, 0.0) + (v.item() if torch.is_tensor(v) else float(v)) * imgs.shape[0]
            n += imgs.shape[0]
        keys = sorted(sums)
        if self.world > 1:
            import torch.distributed as dist

            t = torch.tensor([sums[k] for k in keys] + [float(n)], dtype=torch.float64, device=self.device)
            dist.all_reduce(t)
            vals, n = t[:-1].tolist(), int(t[-1].item())
        else:
            vals = [sums[k] for k in keys]
        out = {k: v / max(n, 1) for k, v in zip(keys, vals)}
        self._log(out, self.global_step)
        return [out]

    def _augmenter(self, train: bool, size: int = 96):
        from .data import GpuAugment

        key = "_aug_train" if train else "_aug_eval"
        if getattr(self, key, None) is None:
            setattr(self, key, GpuAugment(size=size, train=train, seed=73 + 1000 * self.rank))  # rank-distinct boxes
        return getattr(self, key)

    def _prep(self, images: torch.Tensor, train: bool) -> torch.Tensor:
        """uint8 batches (``gpu_augment`` loaders) get the reference transform on the GPU; float batches pass through."""
        if images.dtype != torch.uint8:
            return images
        return self._augmenter(train, images.shape[-1])(images)

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
        if self.world > 1:  # sums over all shards
            tot, correct, nf = [v * self.world for v in self._mean_over_ranks(tot, correct, float(n))]
            n = int(round(nf))
        out = {"val_loss": tot / max(n, 1)}
        if "val_acc" in getattr(module, "logged", {}):
            out["val_acc"] = correct / max(n, 1)
        return out
