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
        for graph, exch in ((False, "nccl"), (True, "nccl"), (True, "peer")):
            m2 = model_registry.create("phoneme_cnn", cfg32).to(dev)
            m2.load_state_dict(nets_oracle.synthetic_state_dict("phoneme_cnn", cfg32, seed=2))
            o2 = FusedClipAdam(m2.parameters(), lr=1e-3)
            t2 = ContrastiveTrainer(m2, [], None, get_loss_fn("supervised_contrastive", temperature=0.15), o2, None, torch.device(dev),
                                    {"gradient_clip_val": 1.0, "progress": False, "cuda_graph": graph, "dp_exchange": exch}, tempfile.mkdtemp(),
                                    logging.getLogger("t"), parallel=ctx)
            m2.train()
            losses = [float(t2.step(x, yl)), float(t2.step(x * 0.5, yl))]
            if graph:
                assert type(t2._graphed).__name__ == ("GraphedDPStepPeer" if exch == "peer" else "GraphedDPStep"), type(t2._graphed)
                if exch == "peer":
                    t2._graphed.check()
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
        (l_e, p_e, s_e, d_e) = o["graph_vs_eager"][0]
        for (l_g, p_g, s_g, d_g) in o["graph_vs_eager"][1:]:          # NCCL segments, then the one-graph peer-memory step
            np.testing.assert_allclose(l_e, l_g, rtol=1e-5)
            n_off = int(((p_e - p_g).abs() > 2e-5).sum())
            assert n_off <= 2e-2 * p_e.numel(), (n_off, p_e.numel())
            assert float((p_e - p_g).abs().max()) <= 2 * 1e-3 * 2
            assert (s_e, d_e, s_g, d_g) == (2, 2, 2, 2)
    for k in (1, 2):
        assert torch.equal(outs[0]["graph_vs_eager"][k][1], outs[1]["graph_vs_eager"][k][1])


# ------------------------------------------------------------------------------------------------ DP step vs R oracle replicas
def _dp_inputs(world, per_rank=16, hw=(40, 50)):
    rs = np.random.RandomState(21)
    xs = [rs.standard_normal((per_rank, 1) + hw).astype(np.float32) for _ in range(world)]
    ys = [np.repeat(np.arange(per_rank // 2) % 5 + 3 * r, 2).astype(np.int64) for r in range(world)]     # classes overlap across ranks
    return xs, ys


def _replica_worker(rank, world, port, out, arch, syncbn=False):
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
        cfg = {"dropout_rate": 0.0} if arch == "phoneme_cnn" else {"dropout_rate": 0.0, "hidden_dims": [64, 64, 128, 128]}
        xs, ys = _dp_inputs(world)
        x, y = torch.from_numpy(xs[rank]).to(dev), torch.from_numpy(ys[rank]).to(dev)
        res = {}
        for mode in ("eager", "graph", "peer"):
            graph = mode != "eager"
            m = model_registry.create(arch, cfg).to(dev)
            m.load_state_dict(nets_oracle.synthetic_state_dict(arch, cfg, seed=6))
            opt = FusedClipAdam(m.parameters(), lr=3e-4, weight_decay=1e-4)
            tr = ContrastiveTrainer(m, [], None, get_loss_fn("supervised_contrastive", temperature=0.15), opt, None, torch.device(dev),
                                    {"gradient_clip_val": 1.0, "progress": False, "cuda_graph": graph, "dp_exchange": "peer" if mode == "peer" else "nccl",
                                     "sync_batchnorm": syncbn},
                                    tempfile.mkdtemp(), logging.getLogger("t"), parallel=ctx)
            m.train()
            assert (m._sync_bn is not None) == bool(syncbn)
            loss = float(tr.step(x, y))
            if syncbn:
                assert m._sync_bn.region.error() == 0
            assert (not graph) or type(tr._graphed).__name__ == ("GraphedDPStepPeer" if mode == "peer" else "GraphedDPStep"), tr._graphed
            if mode == "peer":
                tr._graphed.check()
            names = [n for n, _ in m.named_parameters()]
            res[mode] = dict(loss=loss, grads={n: p.grad.detach().cpu() for n, p in m.named_parameters()},
                                                      params={n: p.detach().cpu() for n, p in m.named_parameters()}, names=names,
                                                      rm=m.state_dict()["projection.1.running_mean"].cpu(),
                                                      buffers={k: v.detach().cpu() for k, v in m.state_dict().items() if "running" in k})
        torch.save(res, os.path.join(out, f"rep{rank}.pt"))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("arch", ["phoneme_cnn", "phoneme_cnn_deep"])
def test_dp_training_step_matches_oracle_replicas(tmp_path, arch):
    """BASELINE configs[4] semantics at a small size (mode (ii)): R ranks with per-rank BatchNorm statistics ==
    R oracle replicas sharing one parameter set whose embeddings are concatenated for ONE global SupCon loss; the all-reduced
    gradient equals the gradient of that loss w.r.t. the shared parameters, and clip + Adam on it gives the ranks' new parameters.
    Eager exchange path, the graphed five-segment step (3 NCCL collectives, split-bucket overlap) and the ONE-graph step with the
    exchanges over NVLink peer memory (csrc/peer.cu) are all checked."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import nets_oracle, optim_oracle, supcon_oracle
    from tests.helpers import analytically_zero_grad
    world, port = 2, _free_port()
    mp.spawn(_replica_worker, args=(world, port, str(tmp_path), arch), nprocs=world, join=True)
    cfg = {"dropout_rate": 0.0} if arch == "phoneme_cnn" else {"dropout_rate": 0.0, "hidden_dims": [64, 64, 128, 128]}
    sd = nets_oracle.synthetic_state_dict(arch, cfg, seed=6)
    xs, ys = _dp_inputs(world)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    embs = []
    for r in range(world):                                   # replica r: its own BatchNorm batch statistics, shared parameters
        live = {k: v.clone() for k, v in sd.items()}
        live.update(params)
        embs.append(nets_oracle.forward(arch, live, torch.from_numpy(xs[r]), training=True))
    loss = supcon_oracle.loss_torch_cpu(torch.cat(embs), torch.from_numpy(np.concatenate(ys)), temperature=0.15)
    loss.backward()
    outs = [torch.load(os.path.join(tmp_path, f"rep{r}.pt")) for r in range(world)]
    names = outs[0]["eager"]["names"]
    gref = [params[n].grad.numpy() for n in names]
    pn = [sd[n].numpy().copy() for n in names]
    optim_oracle.adam_step(pn, gref, [np.zeros_like(p) for p in pn], [np.zeros_like(p) for p in pn], 1, lr=3e-4, weight_decay=1e-4, max_norm=1.0)
    for mode in ("eager", "graph", "peer"):
        for r in range(world):
            o = outs[r][mode]
            assert abs(o["loss"] - float(loss)) <= 1e-4 * abs(float(loss)), (mode, r, o["loss"], float(loss))
            for n, g, pnew in zip(names, gref, pn):
                if analytically_zero_grad(n):
                    continue
                got = o["grads"][n].numpy()
                l2 = np.linalg.norm((got - g).astype(np.float64)) / max(np.linalg.norm(g.astype(np.float64)), 1e-12)
                assert l2 <= 3e-3, (mode, r, n, l2)
                # Adam's first step is lr * sign(g) (|m| / sqrt(v) = 1): elements whose gradient is at round-off level may take
                # either sign, so the update vectors are compared where the reference gradient is not negligible
                du, dr = o["params"][n].numpy() - sd[n].numpy(), pnew - sd[n].numpy()
                sel = np.abs(g) > 0.05 * np.sqrt(np.mean(g.astype(np.float64) ** 2))
                assert np.linalg.norm((du - dr)[sel]) <= 0.02 * np.linalg.norm(dr[sel]) + 1e-9, (mode, r, n)
        for n in names:                                      # every rank holds the same parameters after the step
            assert torch.equal(outs[0][mode]["params"][n], outs[1][mode]["params"][n]), (mode, n)
        assert not torch.equal(outs[0][mode]["rm"], outs[1][mode]["rm"])     # per-rank BatchNorm statistics, as stated


@pytest.mark.timeout(600)
@pytest.mark.parametrize("arch", ["phoneme_cnn", "phoneme_cnn_deep"])
def test_dp_sync_batchnorm_matches_single_process_oracle(tmp_path, arch):
    """SURVEY.md 8e mode (i): with synchronised BatchNorm statistics (peer.SyncStats) R ranks compute what the SINGLE-PROCESS reference
    computes on the concatenated batch -- loss, all-reduced gradients, updated parameters and BatchNorm running statistics -- through
    the eager step, the NCCL-segment graph and the one-graph peer-memory step."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from oracle import nets_oracle, optim_oracle, supcon_oracle
    from tests.helpers import analytically_zero_grad
    world, port = 2, _free_port()
    mp.spawn(_replica_worker, args=(world, port, str(tmp_path), arch, True), nprocs=world, join=True)
    cfg = {"dropout_rate": 0.0} if arch == "phoneme_cnn" else {"dropout_rate": 0.0, "hidden_dims": [64, 64, 128, 128]}
    sd = nets_oracle.synthetic_state_dict(arch, cfg, seed=6)
    xs, ys = _dp_inputs(world)
    params = {k: v.clone().requires_grad_(True) for k, v in sd.items() if v.dtype.is_floating_point and "running" not in k}
    live = {k: v.clone() for k, v in sd.items()}
    live.update(params)
    emb = nets_oracle.forward(arch, live, torch.from_numpy(np.concatenate(xs)), training=True)       # ONE process, the global batch
    loss = supcon_oracle.loss_torch_cpu(emb, torch.from_numpy(np.concatenate(ys)), temperature=0.15)
    loss.backward()
    outs = [torch.load(os.path.join(tmp_path, f"rep{r}.pt")) for r in range(world)]
    names = outs[0]["eager"]["names"]
    gref = [params[n].grad.numpy() for n in names]
    pn = [sd[n].numpy().copy() for n in names]
    optim_oracle.adam_step(pn, gref, [np.zeros_like(p) for p in pn], [np.zeros_like(p) for p in pn], 1, lr=3e-4, weight_decay=1e-4, max_norm=1.0)
    for mode in ("eager", "graph", "peer"):
        for r in range(world):
            o = outs[r][mode]
            assert abs(o["loss"] - float(loss.detach())) <= 1e-4 * abs(float(loss.detach())), (mode, r, o["loss"], float(loss.detach()))
            for n, g, pnew in zip(names, gref, pn):
                if analytically_zero_grad(n):
                    continue
                got = o["grads"][n].numpy()
                l2 = np.linalg.norm((got - g).astype(np.float64)) / max(np.linalg.norm(g.astype(np.float64)), 1e-12)
                assert l2 <= 3e-3, (mode, r, n, l2)
                du, dr = o["params"][n].numpy() - sd[n].numpy(), pnew - sd[n].numpy()
                sel = np.abs(g) > 0.05 * np.sqrt(np.mean(g.astype(np.float64) ** 2))
                assert np.linalg.norm((du - dr)[sel]) <= 0.02 * np.linalg.norm(dr[sel]) + 1e-9, (mode, r, n)
            # running statistics are those of the GLOBAL batch (the oracle's forward updated `live` in place)
            for k, v in o["buffers"].items():
                want = live[k].detach().numpy()
                assert np.abs(v.numpy() - want).max() <= 1e-4 * max(np.abs(want).max(), 1e-3), (mode, r, k)
        for n in names:
            assert torch.equal(outs[0][mode]["params"][n], outs[1][mode]["params"][n]), (mode, n)
        for k in outs[0][mode]["buffers"]:                       # bit-identical on every rank: same sums, same order
            assert torch.equal(outs[0][mode]["buffers"][k], outs[1][mode]["buffers"][k]), (mode, k)
