"""Sharded data parallelism for the hot path (new work: the reference is single-process).

One process per GPU (torchrun), torch.distributed over NCCL/NVLink for plumbing. Per step there are exactly
three small exchanges, all latency-bound at these sizes:

  C1  all_gather of the local embeddings [n,D] and labels [n]  -> every rank sees global negatives;
      rank r then computes ONLY its rows [r*n,(r+1)*n) x all N columns of the SupCon problem (1/R of the work);
  C1' all_gather of the per-row statistics [n,4] -> the backward needs (max, den, n_pos) of REMOTE rows to form
      G_ji, which lets each rank produce dL/dF_local exactly with no gradient reduce-scatter;
  C2  one all-reduce(SUM) of the flat gradient bucket. SUM, not mean: every rank back-propagates the GLOBAL loss
      through its own samples only, so the per-rank parameter gradients are disjoint partial sums.

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
    """Row-block SupCon through libpc_b200.so (pc_supcon_fwd / pc_supcon_bwd)."""

    @staticmethod
    def rows_forward(F, y, temperature, base_temperature, row0, nrows):
        from . import ops
        return ops.supcon_fwd(F, y, None, temperature, base_temperature, row0, nrows)

    @staticmethod
    def rows_backward(F, y, temperature, coef, grad_scale, stats_all, row0, nrows):
        from . import ops
        return ops.supcon_bwd(F, y, None, temperature, coef, grad_scale, stats_all, row0, nrows)


class _ShardedSupCon(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb_local, labels_local, temperature, base_temperature, group, backend):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        n, d = emb_local.shape
        N = n * world
        emb_local = emb_local.contiguous()
        F = torch.empty(N, d, device=emb_local.device, dtype=emb_local.dtype)
        y = torch.empty(N, device=labels_local.device, dtype=torch.int64)
        dist.all_gather_into_tensor(F, emb_local, group=group)                                   # C1
        dist.all_gather_into_tensor(y, labels_local.contiguous().to(torch.int64), group=group)
        row0 = rank * n
        stats, row_loss = backend.rows_forward(F, y, temperature, base_temperature, row0, n)
        stats_all = torch.empty(N, 4, device=F.device, dtype=stats.dtype)
        dist.all_gather_into_tensor(stats_all, stats.contiguous(), group=group)                  # C1'
        total = row_loss.sum(dtype=torch.float32).reshape(1) / N                                 # losses.py:81-82: mean over ALL N rows
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
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
