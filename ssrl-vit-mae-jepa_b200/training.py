"""Host-side mirror of the reference's ``src/training/mae.py`` plus the fused data-parallel step.

``MAEPretrainModule`` keeps the reference constructor ``(model_cfg, training_cfg)``, attributes, and the
``forward / training_step / validation_step / configure_optimizers / on_train_epoch_start`` methods
(reference ``src/training/mae.py:14-83``) and subclasses ``pl.LightningModule`` exactly like the reference: the real
``pytorch_lightning`` when it is installed, else the stand-in under ``shims/`` with the members the module touches
(``save_hyperparameters``, ``log``, ``current_epoch``, ``load_from_checkpoint``).

``FusedTrainer`` is the B200 fast path for the same step: forward + masked MSE + backward + (bucketed
NCCL all-reduce overlapped with the encoder backward) + clip(1.0) + AdamW, all through the C ABI, with no
autograd graph and no host synchronisation.
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Optional, Tuple

import torch
from torch.optim import AdamW
from torch.optim.lr_scheduler import LambdaLR

from ._lib import calls
from .mae import MaskedAutoencoder

from .shims import install as _install_shims

_install_shims()  # appended to sys.path: only found when the real packages are absent
import pytorch_lightning as pl  # noqa: E402  (the real package, or shims/pytorch_lightning)

_Base = pl.LightningModule


def mask_ratio_for_epoch(epoch: int, start: float, end: float, ramp_epochs: int) -> float:
    """reference ``src/training/mae.py:78-82``."""
    progress = min(epoch / max(1, ramp_epochs - 1), 1.0)
    return start + progress * (end - start)


def lr_factor(epoch: int, warmup_epochs: int, total_epochs: int) -> float:
    """reference ``src/training/mae.py:67-70``: linear warm-up x cosine, stepped once per epoch."""
    warmup = (epoch + 1) / max(1, warmup_epochs)
    cosine = 0.5 * (1 + math.cos(math.pi * epoch / total_epochs))
    return min(warmup, 1.0) * cosine


class MAEPretrainModule(_Base):
    """Self-supervised pretraining for Masked Autoencoder (reference ``src/training/mae.py:14``)."""

    def __init__(self, model_cfg: Dict[str, Any], training_cfg: Dict[str, Any]):
        super().__init__()
        self.save_hyperparameters()
        self.model = MaskedAutoencoder(
            general_cfg=model_cfg["general"], encoder_cfg=model_cfg["encoder"], decoder_cfg=model_cfg["decoder"])
        self.mask_start = training_cfg.get("mask_ratio_start", 0.5)
        self.mask_end = training_cfg.get("mask_ratio_end", 0.85)
        self.ramp_epochs = training_cfg.get("mask_ramp_epochs", 200)
        self.lr = float(training_cfg.get("base_learning_rate", 1.5e-4))
        self.weight_decay = float(training_cfg.get("weight_decay", 0.05))
        self.warmup_epochs = int(training_cfg.get("warmup_epochs", 20))
        self.total_epochs = int(training_cfg.get("total_epochs", 200))
        self.batch_size = int(training_cfg.get("batch_size", 512))
        self.criterion = torch.nn.MSELoss()

    def forward(self, x: torch.Tensor):
        return self.model(x)

    def training_step(self, batch, batch_idx):
        imgs, _ = batch
        preds, targets = self(imgs)
        loss = self.criterion(preds.float(), targets)
        self.log("train_loss", loss, prog_bar=True, on_epoch=True)
        return loss

    def validation_step(self, batch, batch_idx):
        imgs, _ = batch
        with torch.no_grad():
            loss, _, _ = self.model.forward_loss(imgs)
        self.log("val_loss", loss, prog_bar=True, on_epoch=True)
        return loss

    @property
    def effective_lr(self) -> float:
        return self.lr * self.batch_size / 256

    def configure_optimizers(self):
        optimizer = AdamW(self.parameters(), lr=self.effective_lr, weight_decay=self.weight_decay)
        scheduler = LambdaLR(optimizer, lambda e: lr_factor(e, self.warmup_epochs, self.total_epochs))
        return {"optimizer": optimizer, "lr_scheduler": {"scheduler": scheduler, "interval": "epoch", "name": "lr"}}

    def on_train_epoch_start(self):
        """Update mask ratio linearly over epochs."""
        new_mask = mask_ratio_for_epoch(self.current_epoch, self.mask_start, self.mask_end, self.ramp_epochs)
        self.model.mask_ratio = new_mask
        self.log("mask_ratio", new_mask, prog_bar=True)


class GradBuckets:
    """Contiguous slices of one flat gradient buffer, all-reduced (averaged) one bucket at a time so a
    bucket's collective can run while the backward pass still produces the next one.  Works on any
    backend (NCCL on the GPUs; gloo in the CPU tests)."""

    def __init__(self, flat: torch.Tensor, boundaries: List[int], group=None):
        assert boundaries[0] == 0 and boundaries[-1] == flat.numel() and sorted(boundaries) == list(boundaries)
        self.flat, self.boundaries, self.group = flat, list(boundaries), group
        self.pending: List[Any] = []

    @property
    def n(self) -> int:
        return len(self.boundaries) - 1

    def launch(self, i: int) -> None:
        import torch.distributed as dist

        lo, hi = self.boundaries[i], self.boundaries[i + 1]
        self._avg = dist.get_backend(self.group) == "nccl"  # gloo has no AVG: sum, then scale in finish()
        if hi > lo:
            op = dist.ReduceOp.AVG if self._avg else dist.ReduceOp.SUM
            self.pending.append(dist.all_reduce(self.flat[lo:hi], op=op, group=self.group, async_op=True))

    def finish(self) -> None:
        import torch.distributed as dist

        for w in self.pending:
            w.wait()
        self.pending.clear()
        if not getattr(self, "_avg", False):
            self.flat.div_(dist.get_world_size(self.group))


class PeerGradReducer:
    """Gradient exchange through NVLink peer memory instead of NCCL (``csrc/collective.cu``): the engine's flat gradient
    buffer is re-homed in symmetric memory (``torch.distributed._symmetric_memory``: one allocation per rank, all of them
    mapped into every process of the node), and per bucket ``mae_allreduce_avg_sumsq`` reads all ranks' slices over
    NVLink, averages them in a fixed order (bit-identical on every rank) into a local buffer and emits the
    sum-of-squares partials of the clip in the same pass.  Cross-rank ordering: a symmetric-memory barrier before each
    bucket (every rank's backward has produced it) and one after the last (nobody still reads what the next backward
    zeroes); everything runs on a side stream, so the compute stream never waits for a peer during backward."""

    NBLK = 128  # CTAs (and sum-of-squares partials) per bucket

    def __init__(self, engine, group=None):
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.engine = engine
        dev = engine.flat.device
        group = group if group is not None else dist.group.WORLD
        if hasattr(symm, "is_symm_mem_enabled_for_group") and not symm.is_symm_mem_enabled_for_group(group.group_name):
            symm.enable_symm_mem_for_group(group.group_name)
        self.sym = symm.empty(engine.n_opt, dtype=torch.float32, device=dev)
        self.hdl = symm.rendezvous(self.sym, group)
        self.sym.zero_()
        engine.grads = self.sym  # backward now accumulates straight into the peer-visible buffer
        self.out = torch.zeros(engine.n_opt, dtype=torch.float32, device=dev)
        self.world = self.hdl.world_size
        self.bounds = engine.grad_bucket_boundaries()
        self.n_partials = self.NBLK * (len(self.bounds) - 1)
        self.stream = torch.cuda.Stream(dev)

    def launch(self, i: int) -> None:
        eng, dev = self.engine, self.engine.flat.device
        lo, hi = self.bounds[i], self.bounds[i + 1]
        cs = self.stream
        cs.wait_stream(torch.cuda.current_stream(dev))
        if eng._side is not None:
            cs.wait_stream(eng._side)
        with torch.cuda.stream(cs):
            self.hdl.barrier(channel=0)  # every rank has produced this bucket
            calls.mae_allreduce_avg_sumsq(self.hdl.buffer_ptrs_dev, self.world, lo, hi - lo, self.out.data_ptr(),
                                          eng.partials.data_ptr() + 4 * self.NBLK * i, self.NBLK, cs.cuda_stream)

    def finish(self) -> None:
        dev = self.engine.flat.device
        with torch.cuda.stream(self.stream):
            self.hdl.barrier(channel=0)  # every rank has read every bucket: the gradient buffers may be reused
        torch.cuda.current_stream(dev).wait_stream(self.stream)


class FusedTrainer:
    """fwd + masked-MSE + bwd + grad all-reduce + clip + AdamW on one stream through the C ABI.

    One process per GPU; parameters are replicated, the batch is sharded, gradients are averaged.  The
    learning rate follows the reference (``effective_lr = base_lr * batch_size / 256`` with the per-process
    batch size from the YAML, ``src/training/mae.py:60``) times the per-epoch warm-up x cosine factor."""

    def __init__(self, module: MAEPretrainModule, distributed: Optional[bool] = None, max_norm: float = 1.0):
        import torch.distributed as dist

        self.module = module
        self.model = module.model
        self.engine = module.model.engine
        self.max_norm = max_norm
        self.epoch = 0
        self.distributed = (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1) \
            if distributed is None else distributed
        self._buckets: Optional[GradBuckets] = None
        self._comm_trigger = None
        # gradient exchange: "p2p" = our own kernel over NVLink peer memory (default on CUDA), "nccl" = bucketed ncclAllReduce
        self.exchange = os.environ.get("MAE_GRAD_EXCHANGE", "p2p")
        self._peer: Optional[PeerGradReducer] = None

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch
        self.module.current_epoch = epoch
        self.module.on_train_epoch_start()

    def current_lr(self) -> float:
        return self.module.effective_lr * lr_factor(self.epoch, self.module.warmup_epochs, self.module.total_epochs)

    def _ensure_buckets(self) -> GradBuckets:
        eng = self.engine
        if self._buckets is None or self._buckets.flat.data_ptr() != eng.grads.data_ptr():
            # one bucket per transformer block (6 for configs/mae.yaml), completed back to front by backward(): only the
            # stem + first encoder block bucket (1.1 MB) is still in flight when backward ends
            self._buckets = GradBuckets(eng.grads, eng.grad_bucket_boundaries())
        return self._buckets

    def _ensure_peer(self) -> Optional[PeerGradReducer]:
        """The peer-memory reducer, or None when NCCL is to be used (``MAE_GRAD_EXCHANGE=nccl``, CPU tensors, or symmetric
        memory cannot be set up on this system: the reason is printed once)."""
        eng = self.engine
        if self.exchange != "p2p" or eng.flat.device.type != "cuda":
            return None
        if self._peer is not None and self._peer.sym.data_ptr() == eng.grads.data_ptr():
            return self._peer
        try:
            self._peer = PeerGradReducer(eng)
        except Exception as e:  # noqa: BLE001 - no P2P / symmetric-memory support here
            import torch.distributed as dist

            if dist.get_rank() == 0:
                print(f"[FusedTrainer] peer-memory gradient exchange unavailable ({e!r}); using NCCL buckets")
            self.exchange, self._peer = "nccl", None
        return self._peer

    def _launch_bucket(self, i: int) -> None:
        """Starts bucket ``i``'s all-reduce behind everything enqueued so far on the compute stream AND on the engine's
        weight-gradient side stream, without making the compute stream wait for the side stream."""
        eng, dev = self.engine, self.engine.flat.device
        if self._comm_trigger is None:
            self._comm_trigger = torch.cuda.Stream(dev) if dev.type == "cuda" else False
        if self._comm_trigger is False:  # CPU tensors (gloo tests)
            self._buckets.launch(i)
            return
        cs = self._comm_trigger
        cs.wait_stream(torch.cuda.current_stream(dev))
        if eng._side is not None:
            cs.wait_stream(eng._side)
        with torch.cuda.stream(cs):
            self._buckets.launch(i)

    def train_step(self, images: torch.Tensor, noise: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One optimisation step; returns the 0-dim device loss tensor (no host sync)."""
        eng = self.engine
        _, _, loss = eng.forward(images, noise, self.model.mask_ratio)
        reduced, n_partials = None, 0
        if self.distributed:
            eng._ensure_flat(images.device)
            peer = self._ensure_peer()
            if peer is not None:
                eng.backward(bucket_cb=peer.launch)
                peer.finish()
                reduced, n_partials = peer.out, peer.n_partials
            else:
                self._ensure_buckets()
                eng.backward(bucket_cb=self._launch_bucket)
                self._buckets.finish()  # the compute stream waits for every bucket's collective
        else:
            eng.backward()
        eng.clip_adamw(self.current_lr(), weight_decay=self.module.weight_decay, max_norm=self.max_norm, grads=reduced,
                       n_partials=n_partials)
        return loss[0].clone()  # the workspace scalar is overwritten by the next step


class DevicePrefetcher:
    """Pinned-host -> device input pipeline: the H2D copy of batch i+1 runs on a side stream while step i
    computes (two device staging buffers).  Iterating yields device tensors that stay valid until the next
    ``next()``.  This is the piece of Lightning's batch transfer the hot path needs (reference
    ``scripts/training/pretrain_mae.py:128`` -> ``trainer.fit`` moves each batch to the GPU before
    ``training_step``), made asynchronous.

    ``transform`` (a ``GpuAugment``; applied to uint8 batches only) runs on the same side stream right behind the copy, writing
    into one of two persistent output buffers: the crop / flip / normalize of batch i+1 then overlaps step i as well,
    instead of sitting at the head of the compute stream."""

    def __init__(self, host_batches, device, transform=None):
        self.it = iter(host_batches)
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.transform = transform
        self.bufs: List[Optional[torch.Tensor]] = [None, None]
        self.obufs: List[Optional[torch.Tensor]] = [None, None]
        self.k = 0
        self.ready: Optional[torch.cuda.Event] = None
        self.cur: Optional[torch.Tensor] = None
        self._preload()

    def _preload(self) -> None:
        try:
            host = next(self.it)
        except StopIteration:
            self.nxt = None
            return
        if isinstance(host, (tuple, list)):
            host = host[0]
        slot = self.k % 2
        self.k += 1
        if self.bufs[slot] is None or self.bufs[slot].shape != host.shape or self.bufs[slot].dtype != host.dtype:
            self.bufs[slot] = torch.empty(host.shape, dtype=host.dtype, device=self.device)
        # everything enqueued so far on the compute stream (incl. the step that last read this slot) goes first
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.bufs[slot].copy_(host, non_blocking=True)
            result = self.bufs[slot]
            if self.transform is not None and host.dtype == torch.uint8:  # float batches are already transformed
                size = getattr(self.transform, "size", host.shape[-1])
                shape = (host.shape[0], host.shape[1], size, size)
                if self.obufs[slot] is None or self.obufs[slot].shape != shape:
                    self.obufs[slot] = torch.empty(shape, dtype=torch.float32, device=self.device)
                result = self.transform(self.bufs[slot], out=self.obufs[slot])
            self.ready = torch.cuda.Event()
            self.ready.record(self.stream)
        self.nxt = result

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        if self.nxt is None:
            raise StopIteration
        torch.cuda.current_stream(self.device).wait_event(self.ready)
        self.cur = self.nxt
        self._preload()
        return self.cur
