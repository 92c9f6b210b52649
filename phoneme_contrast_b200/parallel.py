"""Sharded data parallelism for the hot path (new work: the reference is single-process).

One process per GPU (torchrun), torch.distributed for plumbing. Per step there are exactly three exchanges, all latency-bound at
these sizes (the COUNT of exchanges is what costs):

  C1  ONE gather of the local embeddings with their labels -> every rank sees global negatives; rank r then computes ONLY its rows
      [r*n,(r+1)*n) x all N columns of the SupCon problem (1/R of the work);
  C1' gather of the per-row statistics [n,4] -> the backward needs (max, den, n_pos) of REMOTE rows to form G_ji, which lets each
      rank produce dL/dF_local exactly with no gradient reduce-scatter; the same statistics give every rank the global loss value,
      so there is no all-reduce of the loss;
  C2  all-reduce(SUM) of the flat gradient bucket (the graphed steps split it so that the large late-layer parts overlap the rest
      of the backward). SUM, not mean: every rank back-propagates the GLOBAL loss through its own samples only, so the per-rank
      parameter gradients are disjoint partial sums.

Two transports. The graphed steps (training/graph.py:GraphedDPStepPeer, GraphedShardedLoss) run the exchanges as this library's own
kernels over NVLink peer memory (peer.py, csrc/peer.cu: IPC-mapped regions, flag barriers), so that a whole step is ONE CUDA graph with
no NCCL call on it. The eager path below -- and the fallback when the regions cannot be mapped -- uses NCCL: one all_gather of packed
[n, D+2] rows (label bits in the last two columns), one of the statistics, one all-reduce.

BatchNorm uses per-rank batch statistics (the torch-DDP convention); see DESIGN.md "BatchNorm under DP".
The local row-block compute is pluggable (`backend`) so that the exchange logic is testable on CPU with gloo;
the product backend is the CUDA kernels and nothing else.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
import torch.distributed as dist


class _CudaRowsBackend:
    """Row-block SupCon and the exchange helpers through libpc_b200.so (csrc/supcon*.cu, csrc/dp.cu)."""

    @staticmethod
    def rows_forward(F, y, temperature, base_temperature, row0, nrows):
        from . import ops
        return ops.supcon_fwd(F, y, None, temperature, base_temperature, row0, nrows)

    @staticmethod
    def rows_backward(F, y, temperature, coef, grad_scale, stats_all, row0, nrows):
        from . import ops
        return ops.supcon_bwd(F, y, None, temperature, coef, grad_scale, stats_all, row0, nrows)

    @staticmethod
    def pack(emb, labels):
        from . import ops
        return ops.dp_pack(emb, labels)

    @staticmethod
    def unpack(packed, D):
        from . import ops
        return ops.dp_unpack(packed, D)

    @staticmethod
    def loss_from_stats(stats_all, temperature, base_temperature):
        from . import ops
        return ops.supcon_loss_from_stats(stats_all, temperature, base_temperature, 1.0 / stats_all.shape[0])


class _ShardedSupCon(torch.autograd.Function):
    """Global-batch SupCon over row-sharded embeddings with TWO collectives in the forward and none in the backward:
    all_gather of the packed [n, D+2] (embeddings + label bits) rows, all_gather of the [n, 4] row statistics. The loss value is
    reduced locally, identically on every rank, from the gathered statistics."""

    @staticmethod
    def forward(ctx, emb_local, labels_local, temperature, base_temperature, group, backend):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        n, d = emb_local.shape
        N = n * world
        packed = backend.pack(emb_local.contiguous(), labels_local.contiguous().to(torch.int64))
        packed_all = torch.empty(N, d + 2, device=packed.device, dtype=packed.dtype)
        dist.all_gather_into_tensor(packed_all, packed, group=group)                             # C1: embeddings + labels
        F, y = backend.unpack(packed_all, d)
        row0 = rank * n
        stats, _row_loss = backend.rows_forward(F, y, temperature, base_temperature, row0, n)
        stats_all = torch.empty(N, 4, device=F.device, dtype=stats.dtype)
        dist.all_gather_into_tensor(stats_all, stats.contiguous(), group=group)                  # C1': row statistics
        total = backend.loss_from_stats(stats_all, temperature, base_temperature)                # losses.py:81-82: mean over ALL N rows
        ctx.save_for_backward(F, y, stats_all)
        ctx.cfg = (temperature, base_temperature, row0, n, N, backend)
        return total.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        F, y, stats_all = ctx.saved_tensors
        temperature, base_temperature, row0, n, N, backend = ctx.cfg
        coef = (temperature / base_temperature) / N
        g = grad_out.to(torch.float32).contiguous().view(1)
        dF = backend.rows_backward(F, y, temperature, coef, g, stats_all, row0, n)
        return dF, None, None, None, None, None


class GraphedShardedLoss:
    """Forward + backward of the sharded SupCon loss on static buffers as THREE captured CUDA-graph segments with the two
    all_gathers issued eagerly between them (no autograd, no per-call allocation): at 8192 x 128 the whole problem is 0.29 ms on
    one GPU, so at R ranks the Python / launch overhead of the eager path (about 0.15 ms) is what decides whether sharding pays.

        loss_step = ctx.graphed_loss(loss_fn, emb_local, labels_local)
        loss, dF_local = loss_step(emb_local, labels_local)      # tensors are copied into the static buffers"""

    def __init__(self, ctx: "DataParallelContext", loss_fn, emb_local: torch.Tensor, labels_local: torch.Tensor, warmup: int = 2):
        if getattr(loss_fn, "reduction", "mean") != "mean":
            raise NotImplementedError("the sharded loss implements reduction='mean'")
        self.ctx, self.group = ctx, ctx.group
        backend = ctx.backend
        T = float(loss_fn.temperature)
        Tb = float(getattr(loss_fn, "base_temperature", T))
        n, d = emb_local.shape
        R = ctx.world_size
        N, row0 = n * R, ctx.rank * n
        dev = emb_local.device
        self.emb = emb_local.detach().clone().contiguous()
        self.labels = labels_local.detach().clone().to(torch.int64).contiguous()
        self.packed_all = torch.empty(N, d + 2, device=dev, dtype=torch.float32)
        self.stats_all = torch.empty(N, 4, device=dev, dtype=torch.float32)
        ones = torch.ones(1, device=dev, dtype=torch.float32)

        def seg0():
            self.packed = backend.pack(self.emb, self.labels)

        def seg1():
            self.F, self.y = backend.unpack(self.packed_all, d)
            stats, _ = backend.rows_forward(self.F, self.y, T, Tb, row0, n)
            self.stats = stats.contiguous()

        def seg2():
            self.total = backend.loss_from_stats(self.stats_all, T, Tb)
            self.dF = backend.rows_backward(self.F, self.y, T, (T / Tb) / N, ones, self.stats_all, row0, n)

        self._segs = (seg0, seg1, seg2)
        for _ in range(warmup):
            self._run_eager()
        torch.cuda.synchronize()
        # Exchange over NVLink peer memory (peer.PeerRegion, csrc/peer.cu): the packed rows and the row statistics are stored straight
        # into every rank's gathered buffers by this library's kernels and the whole forward + backward is ONE graph. Falls back
        # (all ranks together) to the three segments with NCCL all_gathers between them when the regions cannot be mapped.
        self.region = None
        if (os.environ.get("PC_DP_EXCHANGE") or "peer") == "peer" and emb_local.is_cuda:
            try:
                self._capture_peer(backend, T, Tb, n, d, N, row0, ones)
            except Exception as exc:      # noqa: BLE001
                import warnings
                warnings.warn(f"sharded loss: peer-memory exchange unavailable ({exc}); using NCCL all_gathers")
                self.region = None
        if self.region is not None:
            return
        pool = torch.cuda.graph_pool_handle()
        self.g = [torch.cuda.CUDAGraph() for _ in range(3)]
        with torch.no_grad():
            for g, seg in zip(self.g, self._segs):
                with torch.cuda.graph(g, pool=pool, capture_error_mode="thread_local"):
                    seg()

    def _capture_peer(self, backend, T, Tb, n, d, N, row0, ones):
        from .peer import PeerRegion
        region = PeerRegion([("F", (N, d), torch.float32), ("y", (N,), torch.int64), ("stats", (N, 4), torch.float32)], self.emb.device,
                            group=self.group)
        stats_all = region.local("stats")
        graph = torch.cuda.CUDAGraph()
        failure = None
        try:
            with torch.no_grad(), torch.cuda.graph(graph, capture_error_mode="thread_local"):
                region.gather_rows(self.emb, self.labels, "F", "y", row0)
                region.barrier(0)
                self.F, self.y = region.local("F"), region.local("y")
                stats, _ = backend.rows_forward(self.F, self.y, T, Tb, row0, n)
                self.stats = stats.contiguous()
                region.bcast(self.stats, "stats", row0 * 16)
                region.barrier(0)
                self.total = backend.loss_from_stats(stats_all, T, Tb)
                self.dF = backend.rows_backward(self.F, self.y, T, (T / Tb) / N, ones, stats_all, row0, n)
        except Exception as exc:      # noqa: BLE001
            failure = f"rank {self.ctx.rank}: {exc}"
        torch.cuda.synchronize()
        outcomes = [None] * self.ctx.world_size
        dist.all_gather_object(outcomes, failure, group=self.group)       # succeed or fall back together; host barrier before the first replay
        outcomes = [o for o in outcomes if o]
        if outcomes:
            region.close()
            raise RuntimeError("; ".join(outcomes))
        self.region, self.graph = region, graph

    def _exchange(self, i):
        if i == 0:
            dist.all_gather_into_tensor(self.packed_all, self.packed, group=self.group)
        elif i == 1:
            dist.all_gather_into_tensor(self.stats_all, self.stats, group=self.group)

    def _run_eager(self):
        with torch.no_grad():
            for i, seg in enumerate(self._segs):
                seg()
                self._exchange(i)

    def __call__(self, emb_local: Optional[torch.Tensor] = None, labels_local: Optional[torch.Tensor] = None):
        if emb_local is not None:
            self.emb.copy_(emb_local, non_blocking=True)
        if labels_local is not None:
            self.labels.copy_(labels_local, non_blocking=True)
        if self.region is not None:
            self.graph.replay()
            return self.total.reshape(()), self.dF
        for i, g in enumerate(self.g):
            g.replay()
            self._exchange(i)
        return self.total.reshape(()), self.dF


class DataParallelContext:
    """Holds the process group and implements the three exchanges. Pass it to ContrastiveTrainer(parallel=...)."""

    def __init__(self, group=None, backend=None):
        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised; call init_distributed() first")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.backend = backend or _CudaRowsBackend

    def loss(self, loss_fn, emb_local: torch.Tensor, labels_local: torch.Tensor) -> torch.Tensor:
        """Global-batch SupCon loss (same value on every rank), differentiable w.r.t. the local embeddings."""
        if getattr(loss_fn, "reduction", "mean") != "mean":
            raise NotImplementedError("the sharded loss implements reduction='mean' (the trainer's configuration)")
        base_t = getattr(loss_fn, "base_temperature", loss_fn.temperature)
        return _ShardedSupCon.apply(emb_local, labels_local, float(loss_fn.temperature), float(base_t), self.group, self.backend)

    def graphed_loss(self, loss_fn, emb_local: torch.Tensor, labels_local: torch.Tensor) -> GraphedShardedLoss:
        """Static-shape, graph-replayed forward + backward of the sharded loss (see GraphedShardedLoss)."""
        return GraphedShardedLoss(self, loss_fn, emb_local, labels_local)

    def all_reduce_gradients(self, flat_grad: torch.Tensor) -> None:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=self.group)                       # C2: one flat bucket

    def all_reduce_parameters(self, model: torch.nn.Module) -> None:
        """Fallback for a stock optimiser: gather grads into one bucket, reduce, scatter back."""
        grads = [p.grad for p in model.parameters() if p.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        self.all_reduce_gradients(flat)
        off = 0
        for g in grads:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()

    def shard(self, global_batch_indices):
        """Contiguous slice of a global batch for this rank (both views of a clip stay together)."""
        n = len(global_batch_indices) // self.world_size
        return global_batch_indices[self.rank * n:(self.rank + 1) * n]

    def broadcast_parameters(self, model: torch.nn.Module, src: int = 0) -> None:
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, src=src, group=self.group)


def init_distributed(backend: Optional[str] = None) -> Optional[DataParallelContext]:
    """Initialise from the torchrun environment (RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*). Returns None when
    WORLD_SIZE <= 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return None
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
    if not dist.is_initialized():
        kw = {}
        if backend == "nccl":
            kw["device_id"] = torch.device("cuda", local_rank)
        dist.init_process_group(backend=backend, **kw)
    return DataParallelContext()
