"""Whole-step CUDA-graph capture of the training hot loop (forward + SupCon + backward + clip + Adam, and the
data-parallel collectives when present). One graph launch replaces ~110 (cnn_small) / ~220 (cnn_deep) kernel
launches and all of the Python sequencing, which is what bounds the step at the reference's batch sizes
("step is latency-bound at reference batch sizes").

Everything that changes from step to step lives in device memory so a replay stays correct: the Adam step count and
learning rate (pc_clip_adam_dev), the Dropout2d call counter (pc_dropout2d_mask's step_dev), BatchNorm running
statistics and num_batches_tracked. Inputs are copied into static buffers before each replay.
"""
from __future__ import annotations

import torch

from .optim import FusedClipAdam


class GraphedTrainStep:
    def __init__(self, trainer, views: torch.Tensor, labels: torch.Tensor, warmup: int = 3):
        if not isinstance(trainer.optimizer, FusedClipAdam):
            raise TypeError("CUDA-graph capture needs the FusedClipAdam optimiser (device-resident step count / lr)")
        self.trainer = trainer
        self.views = views.clone()
        self.labels = labels.clone()
        self.graph = torch.cuda.CUDAGraph()
        opt = trainer.optimizer
        model = trainer.model
        model.train()
        # Warm-up runs real steps (allocator pools, lazy kernel attributes, NCCL channels), so snapshot every piece of
        # training state it touches and put it back: capture must be invisible to the optimisation trajectory.
        snap_opt = (opt.flat_p.clone(), opt.flat_m.clone(), opt.flat_v.clone(), opt._step_dev.clone(), opt._step)
        snap_buf = [b.clone() for b in model.buffers()]
        drop_step = getattr(model, "_drop_step", None)
        snap_drop = drop_step.clone() if drop_step is not None else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                trainer.train_step(self.views, self.labels)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        opt.sync_lr()
        with torch.cuda.graph(self.graph):
            self.loss = trainer.train_step(self.views, self.labels)
        with torch.no_grad():
            opt.flat_p.copy_(snap_opt[0]); opt.flat_m.copy_(snap_opt[1]); opt.flat_v.copy_(snap_opt[2])
            opt._step_dev.copy_(snap_opt[3]); opt._step = snap_opt[4]
            for b, sb in zip(model.buffers(), snap_buf):
                b.copy_(sb)
            if getattr(model, "_drop_step", None) is not None:
                if snap_drop is not None:
                    model._drop_step.copy_(snap_drop)
                else:
                    model._drop_step.zero_()
        self.shape = (tuple(views.shape), tuple(labels.shape))

    def matches(self, views: torch.Tensor, labels: torch.Tensor) -> bool:
        return (tuple(views.shape), tuple(labels.shape)) == self.shape

    def __call__(self, views: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.views.copy_(views, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.trainer.optimizer.sync_lr()
        self.graph.replay()
        self.trainer.optimizer.note_replayed_steps(1)
        return self.loss
