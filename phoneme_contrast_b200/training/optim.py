"""Fused clip_grad_norm_ + Adam on flat buffers: the optimiser half of the hot loop
(src/training/trainer.py:147-152 + torch.optim.Adam from scripts/train.py:129-133), two kernel launches per
step (sum of squares, update) instead of ~10 foreach launches per parameter group.

It is a torch.optim.Optimizer with Adam-compatible state_dict (step / exp_avg / exp_avg_sq per parameter), so
checkpoints written by the trainer (trainer.py:231-245) keep the reference's format. A stock
torch.optim.Adam also still works with the drop-in models; this class is the fast path.
"""
from __future__ import annotations

from typing import Optional

import torch

from .._lib import call, ptr, stream


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 max_grad_norm: Optional[float] = None, grad_prescale: float = 1.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("FusedClipAdam keeps one flat bucket: pass a single parameter group")
        self.max_grad_norm = max_grad_norm
        self.grad_prescale = grad_prescale
        ps = [p for p in self.param_groups[0]["params"] if p.requires_grad]
        if not ps:
            raise ValueError("no trainable parameters")
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedClipAdam runs on CUDA parameters only (no CPU fallback)")
        if any(p.dtype != torch.float32 or p.device != dev for p in ps):
            raise TypeError("all parameters must be float32 on one CUDA device")
        self._params = ps
        self._sizes = [p.numel() for p in ps]
        n = sum(self._sizes)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self._norm_sq = torch.zeros(1, device=dev, dtype=torch.float64)
        self._step = 0
        # device-resident copies of the step count and learning rate: the update kernel reads them from memory, so a
        # step captured in a CUDA graph stays correct across replays (the host refreshes lr_dev when the scheduler moves it)
        self._step_dev = torch.zeros(1, device=dev, dtype=torch.int64)
        self._lr_dev = torch.full((1,), float(lr), device=dev, dtype=torch.float32)
        self._lr_host = float(lr)
        off = 0
        with torch.no_grad():
            for p, k in zip(ps, self._sizes):
                self.flat_p[off:off + k].copy_(p.reshape(-1))
                p.data = self.flat_p[off:off + k].view_as(p)          # parameters become views of the flat bucket
                self.state[p] = {"step": torch.tensor(0.0), "exp_avg": self.flat_m[off:off + k].view_as(p),
                                 "exp_avg_sq": self.flat_v[off:off + k].view_as(p)}
                off += k

    # ---------------------------------------------------------------------------------------------
    def flat_grad(self) -> torch.Tensor:
        """The gradients as ONE flat tensor. Zero-copy when they are views of a single bucket laid out in
        parameter order (what the drop-in models' backward produces); otherwise gathered with one cat."""
        g0 = self._params[0].grad
        if g0 is None:
            raise RuntimeError("FusedClipAdam.step() called before backward()")
        base = g0._base if g0._base is not None else None
        if base is not None and base.dim() == 1 and base.numel() == self.flat_p.numel() and base.is_contiguous():
            off, ok = 0, True
            bp = base.data_ptr()
            for p, k in zip(self._params, self._sizes):
                g = p.grad
                if g is None or g.data_ptr() != bp + 4 * off or not g.is_contiguous():
                    ok = False
                    break
                off += k
            if ok:
                return base
        return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in self._params])

    @torch.no_grad()
    def step(self, closure=None, max_grad_norm: Optional[float] = None, flat_grad: Optional[torch.Tensor] = None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grp = self.param_groups[0]
        g = self.flat_grad() if flat_grad is None else flat_grad
        clip = self.max_grad_norm if max_grad_norm is None else max_grad_norm
        clip = float(clip) if clip else 0.0
        self._step += 1
        n = self.flat_p.numel()
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:
            self.sync_lr()
        call("pc_counter_add", ptr(self._step_dev, torch.int64), 1, stream())
        if clip > 0.0:
            self._norm_sq.zero_()
            call("pc_grad_sumsq", ptr(g), n, ptr(self._norm_sq, torch.float64), stream())
        b1, b2 = grp["betas"]
        call("pc_clip_adam_dev", ptr(self.flat_p), ptr(g), ptr(self.flat_m), ptr(self.flat_v), n, ptr(self._lr_dev), float(b1),
             float(b2), float(grp["eps"]), float(grp["weight_decay"]), clip, ptr(self._norm_sq, torch.float64),
             float(self.grad_prescale), ptr(self._step_dev, torch.int64), stream())
        return loss

    def sync_lr(self) -> None:
        """Push param_groups[0]['lr'] to the device scalar the kernel reads (no-op when unchanged)."""
        lr = float(self.param_groups[0]["lr"])
        if lr != self._lr_host:
            self._lr_dev.fill_(lr)
            self._lr_host = lr

    def note_replayed_steps(self, n: int = 1) -> None:
        """Host-side bookkeeping after CUDA-graph replays (the device counter advanced inside the graph)."""
        self._step += n

    def state_dict(self):
        for p in self._params:
            self.state[p]["step"] = torch.tensor(float(self._step))
        return super().state_dict()

    def total_grad_norm(self) -> torch.Tensor:
        """L2 norm of the (pre-clip, pre-scaled) gradient seen by the last step (device scalar)."""
        return self._norm_sq.sqrt().to(torch.float32) * self.grad_prescale

    def load_state_dict(self, state_dict):
        sd = state_dict["state"]
        groups = state_dict["param_groups"]
        ids = groups[0]["params"]
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        for pid, p in zip(ids, self.param_groups[0]["params"]):
            st = sd.get(pid)
            if st is None or p not in self.state:
                continue
            self.state[p]["exp_avg"].copy_(st["exp_avg"])
            self.state[p]["exp_avg_sq"].copy_(st["exp_avg_sq"])
            self._step = int(float(st["step"]))
            self.state[p]["step"] = torch.tensor(float(self._step))
        self._step_dev.fill_(self._step)
        self._lr_host = None
        self.sync_lr()
