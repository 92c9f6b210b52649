"""2-GPU NCCL test of the data-parallel path (skipped when fewer than two GPUs are visible): the sharded SupCon loss and
the per-rank embedding gradients must equal the single-process result on the concatenated batch, and one DP training
step must leave every rank with identical parameters."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import logging
        import tempfile

        from oracle import nets_oracle
        from phoneme_contrast_b200.models import model_registry
        from phoneme_contrast_b200.parallel import DataParallelContext
        from phoneme_contrast_b200.training import ContrastiveTrainer, FusedClipAdam, get_loss_fn
        ctx = DataParallelContext()
        dev = f"cuda:{rank}"
        rs = np.random.RandomState(0)
        N, D = 256, 128
        f = rs.standard_normal((N, D)).astype(np.float32)
        f /= np.linalg.norm(f, axis=1, keepdims=True)
        y = rs.randint(0, 10, N)
        n = N // world
        fl = torch.from_numpy(f[rank * n:(rank + 1) * n]).to(dev).requires_grad_(True)
        loss = ctx.loss(get_loss_fn("supervised_contrastive", temperature=0.15), fl, torch.from_numpy(y[rank * n:(rank + 1) * n]).to(dev))
        loss.backward()
        res = {"loss": float(loss.detach()), "grad": fl.grad.cpu()}
        # one data-parallel training step (cnn_small, dropout off): parameters must stay identical across ranks
        cfg = {"dropout_rate": 0.0}
        model = model_registry.create("phoneme_cnn", cfg).to(dev)
        model.load_state_dict(nets_oracle.synthetic_state_dict("phoneme_cnn", cfg, seed=2))
        opt = FusedClipAdam(model.parameters(), lr=1e-3)
        tr = ContrastiveTrainer(model, [], None, get_loss_fn("supervised_contrastive", temperature=0.15), opt, None, torch.device(dev),
                                {"gradient_clip_val": 1.0, "progress": False}, tempfile.mkdtemp(), logging.getLogger("t"), parallel=ctx)
        model.train()
        x = torch.from_numpy(np.random.RandomState(10 + rank).standard_normal((8, 1, 40, 50)).astype(np.float32)).to(dev)
        yl = torch.from_numpy(np.repeat(np.arange(4) + 4 * rank, 2)).to(dev)
        res["step_loss"] = float(tr.train_step(x, yl))
        res["param_sum"] = float(opt.flat_p.double().sum())
        res["params"] = opt.flat_p.cpu()
        torch.save(res, os.path.join(out, f"r{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_dp_two_gpus_match_single_process(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import supcon_oracle
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(0)
    N, D = 256, 128
    f = rs.standard_normal((N, D)).astype(np.float32)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    y = rs.randint(0, 10, N)
    want_loss = supcon_oracle.loss(f, y, temperature=0.15)
    want_grad = supcon_oracle.grad(f, y, temperature=0.15)
    outs = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(world)]
    n = N // world
    for r, o in enumerate(outs):
        assert abs(o["loss"] - want_loss) <= 1e-5 * abs(want_loss)
        assert np.abs(o["grad"].numpy() - want_grad[r * n:(r + 1) * n]).max() <= 1e-4 * np.abs(want_grad).max()
    assert outs[0]["step_loss"] == outs[1]["step_loss"]
    assert torch.equal(outs[0]["params"], outs[1]["params"])
