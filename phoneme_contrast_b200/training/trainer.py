"""Drop-in for src/training/trainer.py: ContrastiveTrainer with the reference's constructor, train(),
load_checkpoint(), checkpoint format (trainer.py:231-245) and metrics.json, driving the fused sm_100a path.

What changed inside the hot loop (reference trainer.py:126-164), not in the interface:
  * model / loss / optimiser steps are the fused kernels (no per-op ATen launches);
  * with a FusedClipAdam optimiser, clip_grad_norm_ + Adam.step collapse into two launches;
  * the loss is accumulated on the device and read back once per epoch instead of two .item() syncs per step
    (trainer.py:155,160) -- set config["sync_loss_every_step"]=True to get the reference's per-step postfix;
  * optional data parallelism (phoneme_contrast_b200.parallel): embeddings/labels all_gather for global
    negatives, one flat-bucket gradient all-reduce.
"""
from __future__ import annotations

import json
import logging
from collections import defaultdict
from pathlib import Path
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn
from torch.optim import Optimizer
from torch.utils.data import DataLoader

from .optim import FusedClipAdam

try:  # progress bars are cosmetic
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(it, **kw):
        return it


class ContrastiveTrainer:
    """Trainer for contrastive learning (constructor of reference trainer.py:22-67)."""

    def __init__(self, model: nn.Module, train_loader: DataLoader, val_loader: Optional[DataLoader], loss_fn: nn.Module,
                 optimizer: Optimizer, scheduler, device: torch.device, config: Dict[str, Any], output_dir: Path,
                 logger: logging.Logger, parallel=None):
        self.model = model
        self.train_loader = train_loader
        self.val_loader = val_loader
        self.loss_fn = loss_fn
        self.optimizer = optimizer
        self.scheduler = scheduler
        self.device = torch.device(device)
        self.config = config
        self.output_dir = Path(output_dir)
        self.logger = logger
        self.parallel = parallel          # phoneme_contrast_b200.parallel.DataParallelContext or None
        if self.device.type != "cuda":
            raise RuntimeError("phoneme_contrast_b200 trains on CUDA only (no CPU fallback); got device=%s" % device)
        self.checkpoint_dir = self.output_dir / "checkpoints"
        self.checkpoint_dir.mkdir(parents=True, exist_ok=True)
        self.current_epoch = 0
        self.global_step = 0
        self.best_val_loss = float("inf")
        self.metrics_history = defaultdict(list)
        self._graphed = None            # GraphedTrainStep, built lazily when config["cuda_graph"] is set
        if parallel is not None and config.get("sync_batchnorm") and hasattr(model, "enable_sync_batchnorm") and model._sync_bn is None:
            # BatchNorm statistics over the global batch (SURVEY.md 8e mode (i)); the default is per-rank statistics (mode (ii))
            model.enable_sync_batchnorm(parallel)

    # ------------------------------------------------------------------------------------------- loop
    def train(self, num_epochs: int) -> None:
        self.logger.info(f"Starting training for {num_epochs} epochs")
        self.logger.info(f"Training samples: {len(self.train_loader.dataset)}")
        if self.val_loader:
            self.logger.info(f"Validation samples: {len(self.val_loader.dataset)}")
        for epoch in range(num_epochs):
            self.current_epoch = epoch
            train_metrics = self._train_epoch()
            val_metrics = {}
            if self.val_loader and (epoch + 1) % self.config.get("eval_every", 1) == 0:
                val_metrics = self._validate()
            if self.val_loader and (epoch + 1) % self.config.get("eval_classifier_every", 5) == 0:
                val_metrics.update(self._evaluate_classifier(epoch + 1))
            if self.scheduler:
                self.scheduler.step()
            self._log_metrics(train_metrics, val_metrics)
            if (epoch + 1) % self.config.get("save_every", 10) == 0:
                self._save_checkpoint("periodic")
            best_key = self.config.get("best_metric", "loss")
            metric_for_best = val_metrics.get(best_key, float("inf"))
            if best_key == "loss":
                is_best = metric_for_best < self.best_val_loss
            else:  # accuracy metrics: the reference compares with '>' against the same attribute (trainer.py:111-114)
                is_best = metric_for_best > self.best_val_loss
            if is_best:
                self.best_val_loss = metric_for_best
                self._save_checkpoint("best")
                self.logger.info(f"New best model! {best_key}: {metric_for_best:.4f}")
        self._save_checkpoint("final")
        self._save_metrics()

    def step(self, views: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """train_step, replayed from a captured CUDA graph when config["cuda_graph"] is set and the batch shape is static
        (ContrastiveBatchSampler batches are); falls back to the eager path for odd-shaped batches."""
        if not self.config.get("cuda_graph") or not isinstance(self.optimizer, FusedClipAdam):
            return self.train_step(views, labels)
        if self._graphed is None:
            from .graph import GraphedDPStep, GraphedDPStepPeer, GraphedTrainStep
            kinds = [GraphedTrainStep]
            if self.parallel is not None:
                # data parallel: ONE graph with the exchanges over NVLink peer memory (config["dp_exchange"] = "peer", the default on a
                # CUDA node), or five captured segments with NCCL calls between them ("nccl"; also the fallback when the peer regions
                # cannot be mapped -- every rank falls back together, see peer.PeerRegion)
                import os
                mode = os.environ.get("PC_DP_EXCHANGE") or self.config.get("dp_exchange", "peer")
                kinds = [GraphedDPStepPeer, GraphedDPStep] if mode == "peer" else [GraphedDPStep]
            self._graphed = False
            for kind in kinds:
                try:
                    self._graphed = kind(self, views, labels)
                    break
                except Exception as exc:      # capture is an optimisation: keep training eagerly (training state was restored)
                    self.logger.warning("%s: capture of the training step failed (%s)%s", kind.__name__, exc,
                                        "" if kind is not kinds[-1] else "; continuing with eager launches")
        if self._graphed and self._graphed.matches(views, labels):
            return self._graphed(views, labels)
        return self.train_step(views, labels)

    def train_step(self, views: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """One optimisation step on device-resident inputs; returns the (device) loss. This is the unit bench.py times."""
        embeddings = self._forward_pass(views)
        if self.parallel is not None:
            loss = self.parallel.loss(self.loss_fn, embeddings, labels)
        else:
            loss = self.loss_fn(embeddings, labels)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        clip = self.config.get("gradient_clip_val")
        if isinstance(self.optimizer, FusedClipAdam):
            flat = self.optimizer.flat_grad()
            if self.parallel is not None:
                self.parallel.all_reduce_gradients(flat)          # SUM: each rank holds disjoint partial sums of the global-loss gradient
            self.optimizer.step(max_grad_norm=clip or 0.0, flat_grad=flat)
        else:
            if self.parallel is not None:
                self.parallel.all_reduce_parameters(self.model)
            if clip:
                torch.nn.utils.clip_grad_norm_(self.model.parameters(), clip)
            self.optimizer.step()
        return loss.detach()

    def run_batches(self, batches, on_loss=None) -> Tuple[torch.Tensor, int]:
        """The hot loop of reference trainer.py:138-160 over an iterable of host (or device-resident) batches: one-batch-ahead copy on a
        side stream (`prefetch`), `step`, loss accumulated on the device. Returns (device sum of the losses, number of batches).
        on_loss(i, value): every step's loss is ALSO read on the host, without stalling the pipeline -- step i's scalar is copied to
        pinned memory behind the step and handed to the callback while step i + 1 is already running (the reference reads
        `loss.item()` right after each step, trainer.py:155,160, which drains the GPU every iteration); the last one is delivered
        before this method returns."""
        total = torch.zeros((), device=self.device, dtype=torch.float32)
        n = 0
        pending = None            # (index, pinned scalar, event) of the previous step
        ring = torch.empty(2, dtype=torch.float32).pin_memory() if on_loss is not None else None
        for views, labels in self.prefetch(batches):
            loss = self.step(views, labels)
            total += loss
            if on_loss is not None:
                slot = ring[n & 1:(n & 1) + 1]
                slot.copy_(loss.reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[2].synchronize()
                    on_loss(pending[0], float(pending[1]))
                pending = (n, slot, ev)
            n += 1
            self.global_step += 1
        if pending is not None:
            pending[2].synchronize()
            on_loss(pending[0], float(pending[1]))
        return total, n

    def _train_epoch(self) -> Dict[str, float]:
        self.model.train()
        sync_each = bool(self.config.get("sync_loss_every_step", False))
        pbar = tqdm(self.train_loader, desc=f"Epoch {self.current_epoch + 1}", disable=not self.config.get("progress", True))
        show = (lambda i, v: pbar.set_postfix({"loss": v})) if (sync_each and hasattr(pbar, "set_postfix")) else None
        total, num_batches = self.run_batches(pbar, on_loss=show)
        mean_loss = float(total.item()) / max(num_batches, 1)     # the epoch's only blocking host sync
        self._check_f16_range()
        if self._graphed and hasattr(self._graphed, "check"):
            self._graphed.check()          # peer-memory exchange: a device barrier that timed out (a rank stopped) invalidates the epoch
        return {"loss": mean_loss, "lr": self.optimizer.param_groups[0]["lr"]}

    def _check_f16_range(self) -> None:
        """FP16X2 activation operands are unscaled: if a plane writer saw |a| > 65504 during the epoch (device flag, read here
        next to the loss), the affected steps computed with inf. Switch the model to the range-free tf32x3 engine and say so
        loudly; config["f16_overflow"] = "raise" turns it into an error instead."""
        from .. import _lib as L
        if not hasattr(self.model, "_prec") or not L.f16_overflow(reset=True):
            return
        msg = ("an activation exceeded the fp16 range (65504) on the FP16X2 tensor-core path during this epoch; "
               "the affected steps saw inf operands")
        if self.config.get("f16_overflow", "fallback") == "raise":
            raise FloatingPointError(msg)
        self.logger.error(msg + " -- switching the model to precision 'tf32x3' (no range assumption) from the next step on")
        self.model._prec = L.PREC_TF32X3
        self._graphed = None          # the captured step baked the fp16x2 kernels in

    def _validate(self) -> Dict[str, float]:
        self.model.eval()
        total = torch.zeros((), device=self.device, dtype=torch.float32)
        num_batches = 0
        with torch.no_grad():
            for batch in tqdm(self.val_loader, desc="Validation", disable=not self.config.get("progress", True)):
                views, labels = self._prepare_batch(batch)
                embeddings = self._forward_pass(views)
                total += self.loss_fn(embeddings, labels)
                num_batches += 1
        return {"loss": float(total.item()) / max(num_batches, 1)}

    def prefetch(self, batches):
        """Yields (views, labels) on the device ONE batch ahead: batch i + 1 is copied host -> device on a side stream while step i
        computes (the reference's loop, trainer.py:138-146, copies and computes back to back). Pinned host batches overlap fully."""
        if self.device.type != "cuda":
            for batch in batches:
                yield self._prepare_batch(batch)
            return
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)

        it = iter(batches)

        def fetch():
            # the loader itself runs under the side stream: a device-resident loader (datasets.DeviceFrontendLoader) launches its
            # descriptor upload and front-end kernel there, so they overlap the training step as well
            self._copy_stream.wait_stream(main)        # (buffers freed by the main stream may be reused by the allocator)
            with torch.cuda.stream(self._copy_stream):
                try:
                    batch = next(it)
                except StopIteration:
                    return None
                views, labels = self._prepare_batch(batch)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return views, labels, ev

        nxt = fetch()
        while nxt is not None:
            views, labels, ev = nxt
            nxt = fetch()
            main.wait_event(ev)
            views.record_stream(main)
            labels.record_stream(main)
            yield views, labels

    def _prepare_batch(self, batch: Dict) -> Tuple[torch.Tensor, torch.Tensor]:
        """trainer.py:186-199: H2D, [B,V,C,H,W] -> [B*V,C,H,W], labels repeat_interleave(V)."""
        views = batch["views"].to(self.device, non_blocking=True)
        labels = batch["label"].to(self.device, non_blocking=True)
        if views.dim() == 5:
            b, v = views.shape[:2]
            views = views.reshape(b * v, *views.shape[2:])
            labels = labels.repeat_interleave(v)
        return views, labels

    def _forward_pass(self, views: torch.Tensor) -> torch.Tensor:
        return self.model(views)

    # ------------------------------------------------------------------------------------------- bookkeeping
    def _log_metrics(self, train_metrics: Dict, val_metrics: Dict) -> None:
        for k, v in train_metrics.items():
            self.metrics_history[f"train_{k}"].append(v)
        for k, v in val_metrics.items():
            self.metrics_history[f"val_{k}"].append(v)
        msg = f"Epoch {self.current_epoch + 1} | Train Loss: {train_metrics['loss']:.4f}"
        if "loss" in val_metrics:
            msg += f" | Val Loss: {val_metrics['loss']:.4f}"
        if "linear_accuracy" in val_metrics:
            msg += f" | Linear Acc: {val_metrics['linear_accuracy']:.3f}"
        if "rf_accuracy" in val_metrics:
            msg += f" | RF Acc: {val_metrics['rf_accuracy']:.3f}"
        msg += f" | LR: {train_metrics['lr']:.6f}"
        self.logger.info(msg)

    def _save_checkpoint(self, tag: str) -> None:
        if self.parallel is not None and self.parallel.rank != 0:
            return
        checkpoint = {
            "epoch": self.current_epoch,
            "global_step": self.global_step,
            "model_state_dict": self.model.state_dict(),
            "optimizer_state_dict": self.optimizer.state_dict(),
            "scheduler_state_dict": self.scheduler.state_dict() if self.scheduler else None,
            "best_val_loss": self.best_val_loss,
            "config": self.config,
        }
        path = self.checkpoint_dir / f"checkpoint_{tag}.pt"
        torch.save(checkpoint, path)
        self.logger.info(f"Saved checkpoint: {path}")

    def _save_metrics(self) -> None:
        if self.parallel is not None and self.parallel.rank != 0:
            return
        with open(self.output_dir / "metrics.json", "w") as f:
            json.dump(self.metrics_history, f, indent=2)

    def load_checkpoint(self, path: Path) -> None:
        checkpoint = torch.load(path, map_location=self.device, weights_only=False)
        self.model.load_state_dict(checkpoint["model_state_dict"])
        self.optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
        if self.scheduler and checkpoint["scheduler_state_dict"]:
            self.scheduler.load_state_dict(checkpoint["scheduler_state_dict"])
        self.current_epoch = checkpoint["epoch"]
        self.global_step = checkpoint["global_step"]
        self.best_val_loss = checkpoint["best_val_loss"]
        self.logger.info(f"Loaded checkpoint from epoch {self.current_epoch}")

    def _evaluate_classifier(self, epoch: int) -> Dict[str, float]:
        """Linear / random-forest 5-fold probes on the embeddings (trainer.py:272-323). Host-side sklearn, as in
        the reference: evaluation only, off the hot path; only the embedding extraction runs on the GPU."""
        from sklearn.ensemble import RandomForestClassifier
        from sklearn.linear_model import LogisticRegression
        from sklearn.model_selection import cross_val_score

        self.model.eval()
        embs, labs = [], []
        with torch.no_grad():
            for batch in self.val_loader:
                views, y = self._prepare_batch(batch)
                embs.append(self.model(views).cpu())
                labs.extend(y.cpu().tolist())
            for batch in self.train_loader:
                views, y = self._prepare_batch(batch)
                embs.append(self.model(views).cpu())
                labs.extend(y.cpu().tolist())
        x = torch.cat(embs, dim=0).numpy()
        y = np.array(labs)
        results = {}
        for key, clf in (("linear_accuracy", LogisticRegression(max_iter=1000, random_state=42)),
                         ("rf_accuracy", RandomForestClassifier(n_estimators=100, random_state=42))):
            try:
                results[key] = cross_val_score(clf, x, y, cv=5).mean()
            except (ValueError, RuntimeError) as e:
                self.logger.warning(f"{key} probe failed: {e}")
        return results
