"""The arithmetic behind synchronised BatchNorm statistics in the data-parallel path (SURVEY.md 8e mode (i)), checked on
the CPU against torch autograd on the concatenated batch:

  forward   every rank finalises with the SUM over ranks of (sum y, sum y^2) and the GLOBAL count;
  backward  every rank's apply pass takes the per-rank AVERAGE of the global (sum dz, sum dz xhat) and its LOCAL count M: the
            projection terms are then those of the global batch, and the (dgamma, dbeta) it reports are its share of what the
            gradient all-reduce (SUM over ranks) must deliver.

These are the two substitutions phoneme_contrast_b200.peer.SyncStats makes around unchanged kernels."""
import numpy as np
import torch


def test_sync_batchnorm_substitutions_equal_global_batchnorm():
    torch.manual_seed(0)
    R, M, C, eps = 4, 37, 6, 1e-5            # ranks, rows (pixels) per rank, channels
    y = [torch.randn(M, C, dtype=torch.float64) * (1 + r) + 0.3 * r for r in range(R)]
    gamma = torch.rand(C, dtype=torch.float64) + 0.5
    beta = torch.randn(C, dtype=torch.float64)
    up = [torch.randn(M, C, dtype=torch.float64) for _ in range(R)]          # dL/d(bn output) on each rank

    # the single-process reference on the concatenated batch
    Y = torch.cat(y).requires_grad_(True)
    g, b = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    out = torch.nn.functional.batch_norm(Y, None, None, g, b, training=True, eps=eps)
    out.backward(torch.cat(up))

    # forward substitution: sums over ranks, global count
    s1 = sum(t.sum(0) for t in y)
    s2 = sum((t * t).sum(0) for t in y)
    n = R * M
    mean = s1 / n
    var = s2 / n - mean * mean
    invstd = 1.0 / torch.sqrt(var + eps)
    for r in range(R):
        mine = (y[r] - mean) * invstd * gamma + beta
        np.testing.assert_allclose(mine.numpy(), out.detach()[r * M:(r + 1) * M].numpy(), rtol=1e-10, atol=1e-10)

    # backward substitution: per-rank average of the global sums, LOCAL count
    xhat = [(t - mean) * invstd for t in y]
    sums = [torch.stack([up[r].sum(0), (up[r] * xhat[r]).sum(0)]) for r in range(R)]       # what each rank's reduce pass produces
    avg = sum(sums) / R                                                                     # SyncStats.sync(..., scale = 1 / R)
    dgamma_total, dbeta_total = torch.zeros(C, dtype=torch.float64), torch.zeros(C, dtype=torch.float64)
    for r in range(R):
        dy = gamma * invstd * (up[r] - avg[0] / M - xhat[r] * avg[1] / M)                   # the apply kernel's formula with its local M
        np.testing.assert_allclose(dy.numpy(), Y.grad[r * M:(r + 1) * M].numpy(), rtol=1e-9, atol=1e-12)
        dbeta_total += avg[0]                                                               # what the rank reports ...
        dgamma_total += avg[1]
    np.testing.assert_allclose(dgamma_total.numpy(), g.grad.numpy(), rtol=1e-10)            # ... and the all-reduce sums
    np.testing.assert_allclose(dbeta_total.numpy(), b.grad.numpy(), rtol=1e-10)
