"""Oracle: gradient clipping + Adam step (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, in numpy, what the reference's hot loop does after backward:
  src/training/trainer.py:147-150  torch.nn.utils.clip_grad_norm_(params, gradient_clip_val)
  scripts/train.py:129-133         torch.optim.Adam(lr, weight_decay)  (L2 decay added to the gradient,
                                   i.e. Adam, not AdamW), betas (0.9, 0.999), eps 1e-8
torch semantics followed: clip coefficient = max_norm / (total_norm + 1e-6), clamped to <= 1
(torch/nn/utils/clip_grad.py); Adam single-tensor update (torch/optim/adam.py:_single_tensor_adam).
"""
from __future__ import annotations

import numpy as np


def clip_coef(grads, max_norm: float):
    total = np.sqrt(sum(float((np.asarray(g, dtype=np.float64) ** 2).sum()) for g in grads))
    coef = max_norm / (total + 1e-6)
    return total, min(coef, 1.0)


def adam_step(params, grads, exp_avg, exp_avg_sq, step: int, lr=3e-4, beta1=0.9, beta2=0.999, eps=1e-8,
              weight_decay=1e-4, max_norm=None, dtype=np.float32):
    """One step over lists of arrays; `step` is the 1-based step count AFTER increment.
    Arithmetic is carried in `dtype` (fp32 mirrors torch; fp64 gives a tighter truth). Returns total_norm."""
    total, coef = (0.0, 1.0)
    if max_norm is not None:
        total, coef = clip_coef(grads, max_norm)
    bc1 = 1.0 - beta1 ** step
    bc2 = 1.0 - beta2 ** step
    step_size = lr / bc1
    for p, g, m, v in zip(params, grads, exp_avg, exp_avg_sq):
        g = (np.asarray(g, dtype=dtype) * dtype(coef)).astype(dtype)
        if weight_decay != 0:
            g = (g + dtype(weight_decay) * p).astype(dtype)
        m += (g - m) * dtype(1.0 - beta1)                       # exp_avg.lerp_(grad, 1-beta1)
        v *= dtype(beta2)
        v += dtype(1.0 - beta2) * g * g                         # addcmul_
        denom = (np.sqrt(v) / dtype(np.sqrt(bc2))) + dtype(eps)
        p -= dtype(step_size) * (m / denom)
    return total
