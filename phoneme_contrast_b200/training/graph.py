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
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(warmup):
                    trainer.train_step(self.views, self.labels)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            opt.sync_lr()
            # capture on a HIGH-priority stream: the kernel nodes of the main chain inherit it, the weight-gradient lane's side
            # stream (default = lowest priority) yields the SMs to them whenever both have blocks ready
            import os
            prio = int(os.environ.get("PC_GRAPH_PRIORITY", "0"))      # measured: -1 vs 0 within noise (3.04 - 3.07 ms)
            with torch.cuda.graph(self.graph, stream=torch.cuda.Stream(priority=prio)):
                self.loss = trainer.train_step(self.views, self.labels)
        finally:
            # whether capture succeeded or raised, the warm-up steps must be invisible to the optimisation trajectory (the trainer
            # falls back to eager steps from exactly the state it had before)
            torch.cuda.synchronize()
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


class GraphedDPStep:
    """Data-parallel step as FIVE captured segments with the three NCCL exchanges issued eagerly between them:

        g0 forward -> local embeddings, packed with the label bits        | all_gather(packed rows)                 C1
        g1 unpack, SupCon row block -> row statistics                     | all_gather(row statistics)              C1'
        g2 loss from the gathered statistics, SupCon backward, network backward through the head and the LAST block
                                                                          | all_reduce(bucket tail, async)          C2a
        g3 rest of the network backward                                   | all_reduce(bucket head, async), wait both C2b
        g4 clip + Adam

    The tail of the flat gradient bucket (last block + attention + projection: 74 % of cnn_deep's bytes) is complete when the
    backward is a third of the way through, so its all-reduce runs on NCCL's stream under g3. There is no all-reduce of the
    loss (every rank reduces the gathered row statistics identically) and labels travel inside the embedding gather: 3
    collectives per step instead of round 1's 5. Capturing the collectives themselves inside one whole-step graph deadlocked
    under torchrun in round 1; the segments keep ~150 kernel launches per step inside graphs. The segments talk to the engine
    directly (no autograd): they do exactly what `_NetFunction` / `_ShardedSupCon` do in the eager `train_step`."""

    def __init__(self, trainer, views: torch.Tensor, labels: torch.Tensor, warmup: int = 3):
        import torch.distributed as dist
        from ..models.phoneme_cnn import _prep_input
        if not isinstance(trainer.optimizer, FusedClipAdam):
            raise TypeError("CUDA-graph capture needs the FusedClipAdam optimiser (device-resident step count / lr)")
        par = trainer.parallel
        lf = trainer.loss_fn
        if getattr(lf, "reduction", "mean") != "mean" or not hasattr(lf, "temperature"):
            raise NotImplementedError("the graphed data-parallel step implements the SupCon loss with reduction='mean'")
        self.trainer, self.dist, self.group = trainer, dist, par.group
        opt, model = trainer.optimizer, trainer.model
        model.train()
        dev = views.device
        self.views = views.clone()
        self.labels = labels.clone().to(torch.int64)
        snap_opt = (opt.flat_p.clone(), opt.flat_m.clone(), opt.flat_v.clone(), opt._step_dev.clone(), opt._step)
        snap_buf = [b.clone() for b in model.buffers()]
        drop_step = getattr(model, "_drop_step", None)
        snap_drop = drop_step.clone() if drop_step is not None else None
        for _ in range(warmup):                               # eager steps: allocator pools, NCCL channels, weight-packer recording
            trainer.train_step(self.views, self.labels)
        torch.cuda.synchronize()
        opt.sync_lr()

        n, R = self.views.shape[0], par.world_size
        N, row0 = n * R, par.rank * n
        T = float(lf.temperature)
        Tb = float(getattr(lf, "base_temperature", T))
        clip = float(trainer.config.get("gradient_clip_val") or 0.0)
        backend = par.backend
        params = model._param_list
        self.flat = torch.empty(model._n_param_elems, device=dev, dtype=torch.float32)
        self.split = int(model.tail_bucket_offset())
        grads, off = {}, 0
        for p in params:
            grads[p] = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        ones = torch.ones(1, device=dev, dtype=torch.float32)
        pool = torch.cuda.graph_pool_handle()
        self.g = [torch.cuda.CUDAGraph() for _ in range(5)]
        kw = dict(pool=pool, capture_error_mode="thread_local")
        model._split_backward = True
        try:
            with torch.no_grad():
                with torch.cuda.graph(self.g[0], **kw):
                    emb, saved = model._engine_forward(_prep_input(self.views, model.in_channels), True)
                    self.packed = backend.pack(emb.contiguous(), self.labels)
                d = emb.shape[1]
                self.packed_all = torch.empty(N, d + 2, device=dev, dtype=torch.float32)
                self.stats_all = torch.empty(N, 4, device=dev, dtype=torch.float32)
                with torch.cuda.graph(self.g[1], **kw):
                    F, y = backend.unpack(self.packed_all, d)
                    stats, _ = backend.rows_forward(F, y, T, Tb, row0, n)
                    self.stats = stats.contiguous()
                with torch.cuda.graph(self.g[2], **kw):
                    self.total = backend.loss_from_stats(self.stats_all, T, Tb)
                    dF = backend.rows_backward(F, y, T, (T / Tb) / N, ones, self.stats_all, row0, n)
                    gen = model._engine_backward_gen(saved, dF.contiguous(), grads)
                    next(gen)                                    # head + last block: the bucket tail is complete
                with torch.cuda.graph(self.g[3], **kw):
                    for _ in gen:
                        pass
                with torch.cuda.graph(self.g[4], **kw):
                    opt.step(max_grad_norm=clip, flat_grad=self.flat)
                self._saved = (saved, F, y)
                # the gradients stay visible the usual way: every .grad is a view of the static bucket
                for p in params:
                    if p.requires_grad:
                        p.grad = grads[p]
        finally:
            model._split_backward = False
            # warm-up steps and the host-side counters touched during capture must be invisible to the optimisation trajectory
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
        dist, grp = self.dist, self.group
        self.views.copy_(views, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.trainer.optimizer.sync_lr()
        self.g[0].replay()
        dist.all_gather_into_tensor(self.packed_all, self.packed, group=grp)        # C1  embeddings + labels
        self.g[1].replay()
        dist.all_gather_into_tensor(self.stats_all, self.stats, group=grp)          # C1' row statistics
        self.g[2].replay()
        w_tail = dist.all_reduce(self.flat[self.split:], op=dist.ReduceOp.SUM, group=grp, async_op=True)    # C2a under g3
        self.g[3].replay()
        w_head = dist.all_reduce(self.flat[:self.split], op=dist.ReduceOp.SUM, group=grp, async_op=True)    # C2b
        w_tail.wait()
        w_head.wait()
        self.g[4].replay()
        self.trainer.optimizer.note_replayed_steps(1)
        return self.total.reshape(()).clone()


class GraphedDPStepPeer:
    """Data-parallel step as ONE captured CUDA graph: the three exchanges run over NVLink peer memory in this library's own kernels
    (csrc/peer.cu, phoneme_contrast_b200.peer.PeerRegion) instead of NCCL calls between graph segments.

        forward -> pc_dp_gather_peer stores the local embeddings / labels into EVERY rank's gathered F / y     | barrier       C1
        SupCon row block -> row statistics -> pc_peer_bcast into every rank's statistics buffer             | barrier       C1'
        SupCon backward, network backward; whenever a part of the bucket is complete (head + last block = 74 % of cnn_deep's
        bytes, then the block before it = 18 %):
          side stream: barrier(ch 1), two-shot pc_peer_allreduce of that part (+ the loss value from the gathered statistics)  C2a
          main stream: rest of the network backward; join; barrier; pc_peer_allreduce of the remainder (7 %); barrier        C2b
        clip + Adam

    The flat gradient bucket itself lives in the peer region, so the all-reduce needs no staging copy. One graph launch per step and
    no host work between the exchanges: what is exposed per step is five flag barriers (NVLink round trips) and the head of the
    bucket. The all-reduce adds in rank order and each slice is reduced by exactly one rank, so all ranks hold bit-identical sums."""

    def __init__(self, trainer, views: torch.Tensor, labels: torch.Tensor, warmup: int = 3):
        import os
        import torch.distributed as dist
        from ..models.phoneme_cnn import _prep_input
        from ..peer import PeerRegion
        if not isinstance(trainer.optimizer, FusedClipAdam):
            raise TypeError("CUDA-graph capture needs the FusedClipAdam optimiser (device-resident step count / lr)")
        par = trainer.parallel
        lf = trainer.loss_fn
        if getattr(lf, "reduction", "mean") != "mean" or not hasattr(lf, "temperature"):
            raise NotImplementedError("the graphed data-parallel step implements the SupCon loss with reduction='mean'")
        self.trainer = trainer
        opt, model = trainer.optimizer, trainer.model
        model.train()
        dev = views.device
        n, R = views.shape[0], par.world_size
        N, row0 = n * R, par.rank * n
        d = int(model.embedding_dim)
        n_par = int(model._n_param_elems)
        n_flat = (n_par + 3) // 4 * 4                     # the all-reduce moves 16-byte words; the pad stays zero
        self.region = PeerRegion([("F", (N, d), torch.float32), ("y", (N,), torch.int64), ("stats", (N, 4), torch.float32),
                                  ("flat", (n_flat,), torch.float32)], dev, group=par.group)
        region = self.region
        self.views = views.clone()
        self.labels = labels.clone().to(torch.int64)
        snap_opt = (opt.flat_p.clone(), opt.flat_m.clone(), opt.flat_v.clone(), opt._step_dev.clone(), opt._step)
        snap_buf = [b.clone() for b in model.buffers()]
        drop_step = getattr(model, "_drop_step", None)
        snap_drop = drop_step.clone() if drop_step is not None else None
        for _ in range(warmup):                               # eager steps (NCCL exchanges): allocator pools, weight-packer recording
            trainer.train_step(self.views, self.labels)
        torch.cuda.synchronize()
        opt.sync_lr()

        T = float(lf.temperature)
        Tb = float(getattr(lf, "base_temperature", T))
        clip = float(trainer.config.get("gradient_clip_val") or 0.0)
        backend = par.backend
        params = model._param_list
        self.flat = region.local("flat")[:n_par]
        F, y, stats_all = region.local("F"), region.local("y"), region.local("stats")
        # the bucket is exchanged in up to three parts, each as soon as the backward has completed it (tail_bucket_offsets: last block +
        # head, then the block before it); only the small remainder is reduced after the backward
        cuts = sorted({(int(o) + 3) // 4 * 4 for o in model.tail_bucket_offsets()}, reverse=True)
        ar_blocks = int(os.environ.get("PC_PEER_AR_BLOCKS", "0"))
        grads, off = {}, 0
        for p in params:
            grads[p] = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        ones = torch.ones(1, device=dev, dtype=torch.float32)
        self.graph = torch.cuda.CUDAGraph()
        self._ar_stream = torch.cuda.Stream()
        model._split_backward = "fork"
        failure = None
        try:
            with torch.no_grad():
                with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
                    main = torch.cuda.current_stream()
                    emb, saved = model._engine_forward(_prep_input(self.views, model.in_channels), True)
                    region.gather_rows(emb.contiguous(), self.labels, "F", "y", row0)                     # C1: the gather is the store loop
                    region.barrier(0)
                    stats, _ = backend.rows_forward(F, y, T, Tb, row0, n)
                    region.bcast(stats.contiguous(), "stats", row0 * 16)                                  # C1': row statistics
                    region.barrier(0)
                    dF = backend.rows_backward(F, y, T, (T / Tb) / N, ones, stats_all, row0, n)
                    gen = model._engine_backward_gen(saved, dF.contiguous(), grads)
                    hi, first = n_flat, True
                    for cut in cuts:
                        try:
                            next(gen)                            # the bucket is complete from `cut` on
                        except StopIteration:
                            break
                        if cut >= hi:
                            continue
                        self._ar_stream.wait_stream(main)
                        if getattr(model, "_side_stream", None) is not None:
                            self._ar_stream.wait_stream(model._side_stream)      # this part's weight gradients (not joined into main)
                        with torch.cuda.stream(self._ar_stream):
                            if first:                            # the loss value is off the critical path
                                self.total = backend.loss_from_stats(stats_all, T, Tb)
                                first = False
                            region.barrier(1)
                            region.allreduce("flat", cut, hi - cut, ar_blocks)                            # C2a under the rest of the backward
                        hi = cut
                    for _ in gen:
                        pass
                    if first:
                        self.total = backend.loss_from_stats(stats_all, T, Tb)
                    else:
                        main.wait_stream(self._ar_stream)
                    region.barrier(0)                            # every rank: backward finished, its shares of the earlier parts reduced and stored
                    if hi > 0:
                        region.allreduce("flat", 0, hi, ar_blocks)                                        # C2b: the remainder
                        region.barrier(0)
                    opt.step(max_grad_norm=clip, flat_grad=self.flat)
                self._saved = (saved, F, y, emb, dF, stats)
                for p in params:
                    if p.requires_grad:
                        p.grad = grads[p]
        except Exception as exc:      # noqa: BLE001 -- reported to every rank below
            failure = f"rank {par.rank}: {exc}"
        finally:
            model._split_backward = False
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
        torch.cuda.synchronize()
        # all ranks succeed or fail together (a rank that fell back to NCCL alone would leave the others spinning in a device barrier);
        # the exchange doubles as the host barrier that keeps anyone from replaying while a peer is still capturing
        outcomes = [None] * R
        dist.all_gather_object(outcomes, failure, group=par.group)
        outcomes = [o for o in outcomes if o]
        if outcomes:
            region.close()
            raise RuntimeError("peer-memory capture failed (" + "; ".join(outcomes) + ")")
        self.shape = (tuple(views.shape), tuple(labels.shape))

    def matches(self, views: torch.Tensor, labels: torch.Tensor) -> bool:
        return (tuple(views.shape), tuple(labels.shape)) == self.shape

    def check(self) -> None:
        """Raises when a device barrier of an earlier step timed out (a peer stopped). Synchronises the stream."""
        e = self.region.error(reset=True)
        if e:
            raise RuntimeError(f"data-parallel peer barrier timed out waiting for rank {e - 1}: the steps since the last check are invalid")

    def __call__(self, views: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        self.views.copy_(views, non_blocking=True)
        self.labels.copy_(labels, non_blocking=True)
        self.trainer.optimizer.sync_lr()
        self.graph.replay()
        self.trainer.optimizer.note_replayed_steps(1)
        return self.total.reshape(()).clone()
