"""Drop-in for src/training/losses.py (reference :8-86, :163-173): same class name, constructor, forward
signature, registry function and error behaviour; the arithmetic runs in the fused sm_100a kernels
(csrc/supcon.cu) so the N x N logits never reach HBM.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from .. import ops


class _SupConFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, labels, mask, temperature, base_temperature, reduction):
        feats = features.contiguous()
        if feats.dtype != torch.float32:
            raise TypeError(f"features must be float32, got {feats.dtype}")
        n = feats.shape[0]
        lab = None
        if mask is None:
            lab = labels.contiguous().view(-1).to(device=feats.device, dtype=torch.int64)   # losses.py:53
            if lab.numel() != n:
                raise ValueError(f"labels has {lab.numel()} entries for {n} feature rows")
        else:
            mask = mask.to(device=feats.device, dtype=torch.float32).contiguous()           # losses.py:54 .float()
            if tuple(mask.shape) != (n, n):
                raise ValueError(f"mask must be [{n},{n}], got {tuple(mask.shape)}")
        stats, row_loss = ops.supcon_fwd(feats, lab, mask, temperature, base_temperature)
        if reduction == "mean":
            out = ops.sum_scaled(row_loss, 1.0 / n)
        elif reduction == "sum":
            out = ops.sum_scaled(row_loss, 1.0)
        else:
            out = row_loss                                                                   # losses.py:81-84: no reduction
        ctx.save_for_backward(feats, lab, mask, stats)
        ctx.cfg = (temperature, base_temperature, reduction, n)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        feats, lab, mask, stats = ctx.saved_tensors
        temperature, base_temperature, reduction, n = ctx.cfg
        if reduction not in ("mean", "sum"):
            raise NotImplementedError("backward through reduction=None is not supported by the fused kernel")
        coef = (temperature / base_temperature) / (n if reduction == "mean" else 1)
        g = grad_out.to(torch.float32).contiguous().view(1)
        dF = ops.supcon_bwd(feats, lab, mask, temperature, coef, g, stats)
        return dF, None, None, None, None, None


class SupervisedContrastiveLoss(nn.Module):
    """Supervised Contrastive Loss (Khosla et al., 2020); interface of reference losses.py:8-28."""

    def __init__(self, temperature: float = 0.07, base_temperature: float = 0.07, reduction: str = "mean"):
        super().__init__()
        self.temperature = temperature
        self.base_temperature = base_temperature
        self.reduction = reduction

    def forward(self, features: torch.Tensor, labels: torch.Tensor, mask: Optional[torch.Tensor] = None) -> torch.Tensor:
        if features.shape[0] == 1:
            raise ValueError("Batch size must be greater than 1 for contrastive loss")       # losses.py:44-45
        return _SupConFunction.apply(features, labels, mask, float(self.temperature), float(self.base_temperature),
                                     self.reduction)


class NTXentLoss(nn.Module):
    """With labels this is SupCon without the T/T_base factor (reference losses.py:114-151), i.e. the same
    kernel with base_temperature = temperature. Without labels the reference raises (losses.py:153-159)."""

    def __init__(self, temperature: float = 0.07, reduction: str = "mean"):
        super().__init__()
        self.temperature = temperature
        self.reduction = reduction

    def forward(self, features: torch.Tensor, labels: torch.Tensor = None) -> torch.Tensor:
        if labels is None:
            if features.shape[0] % 2 != 0:
                raise ValueError("Batch size must be even for NT-Xent loss without labels")
            raise NotImplementedError("NT-Xent without labels not implemented in this version")
        return _SupConFunction.apply(features, labels, None, float(self.temperature), float(self.temperature), self.reduction)


_LOSSES = {
    "supervised_contrastive": SupervisedContrastiveLoss,
    "ntxent": NTXentLoss,
}


def get_loss_fn(name: str, **kwargs):
    """Get a loss function by name (reference losses.py:169-173)."""
    if name not in _LOSSES:
        raise ValueError(f"Loss {name} not found. Available: {list(_LOSSES.keys())}")
    return _LOSSES[name](**kwargs)
