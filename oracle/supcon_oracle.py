"""Oracle: supervised contrastive loss (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates src/training/losses.py:26-86 in numpy fp64, plus the analytic gradient
(validated against torch autograd in tests/test_oracle_golden.py),
plus the row-block form used by the data-parallel path.
"""
from __future__ import annotations

import numpy as np


def _positive_mask(labels, mask, n):
    if mask is None:
        y = np.asarray(labels).reshape(-1, 1)
        P = (y == y.T).astype(np.float64)                      # losses.py:53-54
    else:
        P = np.asarray(mask, dtype=np.float64)
    O = 1.0 - np.eye(n)                                         # losses.py:57
    return P * O, O                                             # losses.py:58


def row_stats(features, labels=None, mask=None, temperature=0.07, rows=None):
    """Per-row quantities for rows `rows` (slice) against all N columns.

    Returns dict(m, den, npos, spos): row max of z (incl. diagonal, losses.py:64), denominator
    sum_j exp(z-m) O + 1e-6 (losses.py:68-69), number of positives (losses.py:73) and
    sum_j P (z - m)."""
    F = np.asarray(features, dtype=np.float64)
    n = F.shape[0]
    rows = slice(0, n) if rows is None else rows
    P, O = _positive_mask(labels, mask, n)
    z = (F[rows] @ F.T) / temperature                           # losses.py:49,61
    m = z.max(axis=1)
    zc = z - m[:, None]
    den = (np.exp(zc) * O[rows]).sum(1) + 1e-6
    npos = P[rows].sum(1)
    spos = (P[rows] * zc).sum(1)
    return {"m": m, "den": den, "npos": npos, "spos": spos}


def loss(features, labels=None, mask=None, temperature=0.07, base_temperature=0.07, reduction="mean"):
    """SupervisedContrastiveLoss.forward (losses.py:26-86)."""
    F = np.asarray(features, dtype=np.float64)
    n = F.shape[0]
    if n == 1:
        raise ValueError("Batch size must be greater than 1 for contrastive loss")   # losses.py:44-45
    st = row_stats(F, labels, mask, temperature)
    nn_ = np.where(st["npos"] == 0, 1.0, st["npos"])                                   # losses.py:73-74
    mean_log_prob_pos = (st["spos"] - st["npos"] * np.log(st["den"])) / nn_            # losses.py:69,76
    per_row = -(temperature / base_temperature) * mean_log_prob_pos                    # losses.py:79
    if reduction == "mean":
        return per_row.mean()
    if reduction == "sum":
        return per_row.sum()
    return per_row


def grad(features, labels=None, mask=None, temperature=0.07, base_temperature=0.07, reduction="mean",
         grad_out=1.0, rows=None):
    """dL/dF. With `rows`, only the gradient of the full loss w.r.t. F[rows] is returned, computed the way
    the row-sharded DP path does: it needs only row stats of all rows (an all_gather of [N,3]),
    no gradient reduce-scatter."""
    F = np.asarray(features, dtype=np.float64)
    n = F.shape[0]
    P, O = _positive_mask(labels, mask, n)
    st = row_stats(F, labels, mask, temperature)
    z = (F @ F.T) / temperature
    e = np.exp(z - st["m"][:, None]) * O
    nn_ = np.where(st["npos"] == 0, 1.0, st["npos"])
    h = (st["npos"] > 0).astype(np.float64) if mask is None else (st["npos"] != 0).astype(np.float64)
    c = (temperature / base_temperature) * grad_out
    if reduction == "mean":
        c = c / n
    # d loss / d z_ij  (row max is detached, losses.py:65)
    # note: with a float user mask, d/dz of -(1/n_i) sum_j P_ij (z_ij - log den_i) = -(P_ij/n_i - (sum_j P_ij / n_i) e_ij/den_i)
    w = (P.sum(1) / nn_) if mask is not None else h
    G = -c * (P / nn_[:, None] - w[:, None] * e / st["den"][:, None])
    full = (G + G.T) @ F / temperature
    return full if rows is None else full[rows]


def loss_torch_cpu(features, labels, temperature=0.15, base_temperature=0.07):
    """fp32 torch restatement with the reference's op sequence (used for CPU timing and autograd checks)."""
    import torch

    n = features.shape[0]
    sim = torch.matmul(features, features.T)
    y = labels.contiguous().view(-1, 1)
    mask = torch.eq(y, y.T).float()
    lm = torch.ones_like(mask) - torch.eye(n)
    mask = mask * lm
    logits = sim / temperature
    logits = logits - logits.max(dim=1, keepdim=True)[0].detach()
    exp_logits = torch.exp(logits) * lm
    log_prob = logits - torch.log(exp_logits.sum(1, keepdim=True) + 1e-6)
    ms = mask.sum(1)
    ms = torch.where(ms == 0, torch.ones_like(ms), ms)
    mlpp = (mask * log_prob).sum(1) / ms
    return (-(temperature / base_temperature) * mlpp).mean()
