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
        # the same two steps through the graphed data-parallel step (four captured segments + eager NCCL) and eagerly:
        # exact-fp32 kernels so both trajectories are deterministic
        cfg32 = {"dropout_rate": 0.0, "precision": "fp32"}
        finals = []
        for graph in (False, True):
            m2 = model_registry.create("phoneme_cnn", cfg32).to(dev)
            m2.load_state_dict(nets_oracle.synthetic_state_dict("phoneme_cnn", cfg32, seed=2))
            o2 = FusedClipAdam(m2.parameters(), lr=1e-3)
            t2 = ContrastiveTrainer(m2, [], None, get_loss_fn("supervised_contrastive", temperature=0.15), o2, None, torch.device(dev),
                                    {"gradient_clip_val": 1.0, "progress": False, "cuda_graph": graph}, tempfile.mkdtemp(),
                                    logging.getLogger("t"), parallel=ctx)
            m2.train()
            losses = [float(t2.step(x, yl)), float(t2.step(x * 0.5, yl))]
            finals.append((losses, o2.flat_p.cpu(), o2._step, int(o2._step_dev.item())))
        res["graph_vs_eager"] = finals
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
    # graphed data-parallel step == eager data-parallel step (same losses; parameters identical up to the order-dependent last
    # bit of the fp64 BatchNorm statistics, cf. test_cuda_graph_step_matches_eager), on every rank, step counters advanced
    for o in outs:
        (l_e, p_e, s_e, d_e), (l_g, p_g, s_g, d_g) = o["graph_vs_eager"]
        np.testing.assert_allclose(l_e, l_g, rtol=1e-5)
        n_off = int(((p_e - p_g).abs() > 2e-5).sum())
        assert n_off <= 2e-2 * p_e.numel(), (n_off, p_e.numel())
        assert float((p_e - p_g).abs().max()) <= 2 * 1e-3 * 2
        assert (s_e, d_e, s_g, d_g) == (2, 2, 2, 2)
    assert torch.equal(outs[0]["graph_vs_eager"][1][1], outs[1]["graph_vs_eager"][1][1])
