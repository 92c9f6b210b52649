"""World-size-2 gloo test (CPU) of the data-parallel exchange logic in phoneme_contrast_b200/parallel.py.
The local row-block compute is injected from the oracle (tests may use the oracle as the checker); what is under
test is the sharding, the three collectives and the gradient assembly -- the product backend is the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


from tests.exchange_torch import TorchExchangeMixin


class OracleRowsBackend(TorchExchangeMixin):
    @staticmethod
    def rows_forward(F, y, temperature, base_temperature, row0, nrows):
        from oracle import supcon_oracle
        rows = slice(row0, row0 + nrows)
        st = supcon_oracle.row_stats(F.numpy(), y.numpy(), temperature=temperature, rows=rows)
        nn_ = np.where(st["npos"] == 0, 1.0, st["npos"])
        row_loss = -(temperature / base_temperature) * (st["spos"] - st["npos"] * np.log(st["den"])) / nn_
        stats = np.stack([st["m"], st["den"], st["npos"], st["spos"]], 1)
        return torch.from_numpy(stats).to(F.dtype), torch.from_numpy(row_loss).to(F.dtype)

    @staticmethod
    def rows_backward(F, y, temperature, coef, grad_scale, stats_all, row0, nrows):
        from oracle import supcon_oracle
        n = F.shape[0]
        base_t = temperature / (coef * n)
        g = supcon_oracle.grad(F.numpy(), y.numpy(), temperature=temperature, base_temperature=base_t,
                               grad_out=float(grad_scale[0]), rows=slice(row0, row0 + nrows))
        return torch.from_numpy(g).to(F.dtype)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from phoneme_contrast_b200.parallel import DataParallelContext
        from phoneme_contrast_b200.training.losses import SupervisedContrastiveLoss
        ctx = DataParallelContext(backend=OracleRowsBackend)
        rs = np.random.RandomState(0)
        N, D = 48, 32
        f = rs.standard_normal((N, D))
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        y = rs.randint(0, 6, N)
        n = N // world
        # a linear "model" shared by the ranks so that the parameter-gradient all-reduce can be checked too
        w = torch.eye(D, dtype=torch.float64).requires_grad_(True)
        emb_local = torch.from_numpy(f[rank * n:(rank + 1) * n]) @ w
        loss = ctx.loss(SupervisedContrastiveLoss(temperature=0.15), emb_local, torch.from_numpy(y[rank * n:(rank + 1) * n]))
        loss.backward()
        flat = w.grad.reshape(-1).clone()
        ctx.all_reduce_gradients(flat)
        assert ctx.shard(list(range(10))) == list(range(10))[rank * 5:(rank + 1) * 5]
        torch.save({"loss": float(loss.detach()), "dw": flat.reshape(D, D)}, os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_sharded_supcon_world2_gloo(tmp_path):
    from oracle import supcon_oracle
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(0)
    N, D = 48, 32
    f = rs.standard_normal((N, D))
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    y = rs.randint(0, 6, N)
    want_loss = supcon_oracle.loss(f, y, temperature=0.15)
    want_dw = f.T @ supcon_oracle.grad(f, y, temperature=0.15)            # d loss / d w for emb = f @ w at w = I
    for r in range(world):
        got = torch.load(os.path.join(tmp_path, f"r{r}.pt"))
        assert abs(got["loss"] - want_loss) < 1e-6 * abs(want_loss)   # the product reduces the row losses in fp32
        np.testing.assert_allclose(got["dw"].numpy(), want_dw, rtol=1e-9, atol=1e-12)
