"""Peer-memory exchange kernels (csrc/peer.cu) on ONE device: R "ranks" are R regions of the same GPU driven from R streams, so the
pack / broadcast / all-reduce / barrier logic is covered on the single-GPU box too (the IPC mapping itself and the whole-step graph
need two GPUs: tests/test_gpu_multi.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _regions(R, fields, timeout_ms=5000):
    from phoneme_contrast_b200.peer import PeerRegion
    probe = PeerRegion(fields, "cuda", bases=[16] * R, rank=0, buf=None)
    bufs = [torch.zeros(probe.nbytes, device="cuda", dtype=torch.uint8) for _ in range(R)]
    bases = [b.data_ptr() for b in bufs]
    return [PeerRegion(fields, "cuda", bases=bases, rank=r, buf=bufs[r], timeout_ms=timeout_ms) for r in range(R)]


@pytest.mark.parametrize("R", [2, 3, 4])
def test_peer_exchanges_single_device(R):
    n, D, P = 24, 128, 4 * 12345
    N = n * R
    regs = _regions(R, [("F", (N, D), torch.float32), ("y", (N,), torch.int64), ("stats", (N, 4), torch.float32), ("flat", (P,), torch.float32)])
    rs = np.random.RandomState(R)
    emb = [torch.from_numpy(rs.standard_normal((n, D)).astype(np.float32)).cuda() for _ in range(R)]
    lab = [torch.from_numpy(rs.randint(-2 ** 40, 2 ** 40, n)).cuda() for _ in range(R)]
    st = [torch.from_numpy(rs.standard_normal((n, 4)).astype(np.float32)).cuda() for _ in range(R)]
    g = [torch.from_numpy(rs.standard_normal(P).astype(np.float32)).cuda() for _ in range(R)]
    streams = [torch.cuda.Stream() for _ in range(R)]
    split = 4 * 3000
    # every kernel once before the first cross-"rank" wait: the first launch of a kernel loads it (CUDA lazy loading), and that load can
    # block behind a barrier kernel that is spinning on this same device (in production each rank has its own device and the
    # kernels are loaded when the step is captured)
    solo = _regions(1, [("x", (4,), torch.float32)])[0]
    solo.barrier(0)
    solo.barrier(1)
    regs[0].gather_rows(emb[0], lab[0], "F", "y", 0)
    regs[0].bcast(st[0], "stats", 0)
    regs[0].allreduce("flat", split, P - split)
    torch.cuda.synchronize()
    for rep in range(3):                       # several rounds: epochs advance, buffers are reused
        for r in range(R):
            regs[r].local("flat").copy_(g[r] * (rep + 1))
        torch.cuda.synchronize()
        for r in range(R):
            with torch.cuda.stream(streams[r]):
                regs[r].gather_rows(emb[r], lab[r], "F", "y", r * n)
                regs[r].bcast(st[r], "stats", r * n * 16)
                regs[r].barrier(0)
                regs[r].barrier(1)
                regs[r].allreduce("flat", split, P - split)
                regs[r].barrier(1)
                regs[r].allreduce("flat", 0, split, blocks=3)
                regs[r].barrier(0)
        torch.cuda.synchronize()
        want = torch.stack(g).double().sum(0) * (rep + 1)
        errs = [regs[r].error() for r in range(R)]
        if any(errs):
            # a barrier timed out: the R barrier kernels (one warp each, on R streams) were not co-scheduled -- e.g. under a tool that
            # serialises kernel launches. The exchange logic then cannot be exercised on one device; the 2-GPU tests cover it.
            pytest.skip(f"kernels of different streams did not run concurrently on this device (barrier errors {errs})")
        for r in range(R):
            assert torch.equal(regs[r].local("F"), torch.cat(emb)) and torch.equal(regs[r].local("y"), torch.cat(lab))
            assert torch.equal(regs[r].local("stats"), torch.cat(st))
            got = regs[r].local("flat")
            assert torch.equal(got, regs[0].local("flat"))                # bit-identical on every rank
            assert float((got.double() - want).abs().max()) <= 1e-5 * float(want.abs().max())
    if R == 2:                                 # two addends: the sum is exact in any order
        assert torch.equal(regs[0].local("flat"), (g[0] * 3 + g[1] * 3))


def test_peer_barrier_timeout_sets_error():
    regs = _regions(2, [("x", (4,), torch.float32)], timeout_ms=50)
    regs[0].barrier(0)                         # rank 1 never arrives
    torch.cuda.synchronize()
    assert regs[0].error(reset=False) == 2
    regs[0].barrier(0)                         # sticky: returns at once
    torch.cuda.synchronize()
    assert regs[0].error(reset=True) == 2
    assert regs[0].error() == 0


def test_peer_export_reports_offset():
    import ctypes as C
    from phoneme_contrast_b200 import _lib as L
    big = torch.zeros(1 << 20, device="cuda", dtype=torch.uint8)
    view = big[4096:]
    h1, h2 = (C.c_ubyte * 64)(), (C.c_ubyte * 64)()
    o1, o2 = C.c_size_t(0), C.c_size_t(0)
    L.check(L.lib().pc_peer_export(C.c_void_p(big.data_ptr()), h1, C.byref(o1)))
    L.check(L.lib().pc_peer_export(C.c_void_p(view.data_ptr()), h2, C.byref(o2)))
    assert bytes(h1) == bytes(h2) and o2.value - o1.value == 4096
