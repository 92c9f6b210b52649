"""Peer-memory regions for the data-parallel exchanges (csrc/peer.cu): one device buffer per rank, laid out identically on all
ranks of the node and mapped into every rank's address space (CUDA IPC over NVLink / NVSwitch), so that the embedding / statistics
all_gathers and the gradient all-reduce are stores and loads of this library's own kernels -- stream-ordered, graph-capturable, with
no NCCL call on the step (SURVEY.md 8e C1 / C1' / C2; the single-process reference has no counterpart).

torch.distributed is used ONCE, at construction, to exchange the 64-byte IPC handles (all_gather_object) and to line the ranks up
(barrier); torch owns the memory (the region is an ordinary uint8 tensor of the caching allocator, carved into typed views).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import torch

from . import _lib as L


# IPC handles name whole cudaMalloc blocks, and the caching allocator places several tensors in one block: two regions of a peer can
# arrive with the SAME handle. A handle is mapped once per process and the mapping is kept until exit (torch's own IPC cache does the same).
_MAPPED: dict = {}


class _OwnBlock:
    """A cudaMalloc block of the library's own, visible to torch through __cuda_array_interface__ (zero-copy uint8 view)."""

    def __init__(self, nbytes: int):
        base = C.c_void_p()
        L.check(L.lib().pc_peer_alloc(nbytes, C.byref(base)))
        self.ptr, self.nbytes = int(base.value), int(nbytes)
        self.__cuda_array_interface__ = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2, "strides": None}

    def __del__(self):
        try:
            L.lib().pc_peer_free(C.c_void_p(self.ptr))
        except Exception:      # noqa: BLE001 -- interpreter shutdown
            pass


def _align(x: int, a: int = 256) -> int:
    return (x + a - 1) // a * a


class PeerRegion:
    """`fields`: [(name, shape, dtype)], allocated in this order behind the flag block. `local(name)` is this rank's typed view,
    `offset(name)` its byte offset inside every region."""

    def __init__(self, fields: Sequence[Tuple[str, Tuple[int, ...], torch.dtype]], device, group=None, bases: List[int] = None, rank: int = None,
                 buf: torch.Tensor = None, timeout_ms: int = 20000):
        lib = L.lib()
        self.flag_off = 0
        off = _align(lib.pc_peer_flag_bytes())
        self._fields = {}
        for name, shape, dtype in fields:
            nbytes = int(torch.empty((), dtype=dtype).element_size())
            for s in shape:
                nbytes *= int(s)
            self._fields[name] = (off, tuple(int(s) for s in shape), dtype)
            off = _align(off + nbytes)
        self.nbytes = off
        self.timeout_ms = int(timeout_ms)
        if bases is not None:
            # single-process form (tests: several "ranks" on one device): the caller supplies the regions' buffers / addresses
            self.buf, self.rank, self.world_size = buf, int(rank), len(bases)
            self._bases_py = list(bases)
        else:
            import torch.distributed as dist
            self.rank, self.world_size = dist.get_rank(group), dist.get_world_size(group)
            if self.world_size > lib.pc_peer_max_ranks():
                raise NotImplementedError(f"peer regions support up to {lib.pc_peer_max_ranks()} ranks of one node")
            self.buf = torch.zeros(self.nbytes, device=device, dtype=torch.uint8)
            torch.cuda.synchronize(device)
            # export; a rank that cannot (allocator without cudaMalloc blocks, IPC forbidden in the container) says so to everyone, so
            # that all ranks raise together and the caller falls back to the NCCL exchanges on every rank alike
            mine = None
            handle = (C.c_ubyte * 64)()
            offset = C.c_size_t(0)
            try:
                L.check(lib.pc_peer_export(C.c_void_p(self.buf.data_ptr()), handle, C.byref(offset)))
                mine = (bytes(handle), int(offset.value))
            except Exception as exc:      # noqa: BLE001
                # torch's allocator handed out memory that CUDA IPC cannot export (expandable segments): take a cudaMalloc block of the
                # library's own instead and view it as a tensor
                try:
                    self._own = _OwnBlock(self.nbytes)
                    self.buf = torch.as_tensor(self._own, device=device)
                    L.check(lib.pc_peer_export(C.c_void_p(self.buf.data_ptr()), handle, C.byref(offset)))
                    mine = (bytes(handle), int(offset.value))
                except Exception as exc2:      # noqa: BLE001
                    mine = ("error", f"{exc}; own block: {exc2}")
            everyone = [None] * self.world_size
            dist.all_gather_object(everyone, mine, group=group)
            bad = [f"rank {r}: {e[1]}" for r, e in enumerate(everyone) if e[0] == "error"]
            if bad:
                raise RuntimeError("peer region export failed (" + "; ".join(bad) + ")")
            self._bases_py, fail = [], None
            for r, (h, o) in enumerate(everyone):
                if r == self.rank:
                    self._bases_py.append(self.buf.data_ptr())
                    continue
                if h not in _MAPPED:
                    base = C.c_void_p()
                    hb = (C.c_ubyte * 64).from_buffer_copy(h)
                    rc = lib.pc_peer_open(hb, C.byref(base))
                    if rc != 0:
                        fail = f"rank {self.rank} could not map rank {r}: {L.last_error()}"
                        break
                    _MAPPED[h] = base.value
                self._bases_py.append(_MAPPED[h] + o)
            fails = [None] * self.world_size
            dist.all_gather_object(fails, fail, group=group)     # doubles as the barrier: everyone zeroed its flags and mapped everybody
            fails = [f for f in fails if f]
            if fails:
                self.close()
                raise RuntimeError("peer region mapping failed (" + "; ".join(fails) + ")")
        self.bases = (C.c_ulonglong * self.world_size)(*self._bases_py)

    # ---------------------------------------------------------------------------------------------------------- layout
    def offset(self, name: str) -> int:
        return self._fields[name][0]

    def local(self, name: str) -> torch.Tensor:
        off, shape, dtype = self._fields[name]
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        return self.buf[off:off + nbytes].view(dtype).view(shape)

    # ---------------------------------------------------------------------------------------------------------- exchanges
    def barrier(self, channel: int = 0) -> None:
        L.call("pc_peer_barrier", self.bases, self.world_size, self.rank, self.flag_off, int(channel), self.timeout_ms, L.stream())

    def gather_rows(self, emb: torch.Tensor, labels: torch.Tensor, f_name: str, y_name: str, row0: int) -> None:
        """emb [n, D] -> rows [row0, row0 + n) of the [N, D] field f_name, labels [n] int64 -> the same rows of the [N] field y_name,
        of EVERY rank."""
        n, D = emb.shape
        L.call("pc_dp_gather_peer", L.ptr(emb), L.ptr(labels, torch.int64), n, D, self.bases, self.world_size, self.offset(f_name),
               self.offset(y_name), int(row0), L.stream())

    def bcast(self, src: torch.Tensor, name: str, byte_off: int = 0) -> None:
        """The bytes of `src` -> byte offset byte_off of the field `name` of EVERY rank."""
        nbytes = src.numel() * src.element_size()
        L.call("pc_peer_bcast", L.ptr(src, None), nbytes, self.bases, self.world_size, self.offset(name) + int(byte_off), L.stream())

    def allreduce(self, name: str, start: int = 0, count: int = None, blocks: int = 0) -> None:
        """In-place sum over the ranks of elements [start, start + count) of the fp32 field `name` (bracket with barriers)."""
        off, shape, dtype = self._fields[name]
        assert dtype == torch.float32
        total = 1
        for s in shape:
            total *= s
        count = total - start if count is None else count
        L.call("pc_peer_allreduce", self.bases, self.world_size, self.rank, off + 4 * int(start), int(count), int(blocks), L.stream())

    def error(self, reset: bool = True) -> int:
        """0, or 1 + the rank a barrier timed out on (sticky). Synchronises the current stream."""
        out = C.c_int(0)
        L.check(L.lib().pc_peer_error(self.bases, self.world_size, self.rank, self.flag_off, 1 if reset else 0, C.byref(out), L.stream()))
        return int(out.value)

    def close(self) -> None:
        """Drops this region's references; the peer mappings themselves stay cached for the life of the process (_MAPPED)."""
        self._bases_py = []


class SyncStats:
    """Synchronised BatchNorm statistics (SURVEY.md 8e mode (i), C3): a one-shot all-reduce of small fp64 blocks over peer memory.
    Every call site of a step takes the next slice of the region's slot array (same order on every rank; begin_step() rewinds), stores
    its block into slot `rank` of EVERY region, and after ONE flag barrier each rank sums its local slots in rank order -- the results
    are bit-identical on all ranks. Stream-ordered kernels only, so the exchanges work in eager steps and inside captured graphs."""

    def __init__(self, device, group=None, capacity: int = 1 << 15, region: PeerRegion = None, world_size: int = None):
        if region is None:
            import torch.distributed as dist
            world_size = dist.get_world_size(group)
            region = PeerRegion([("bn", (world_size * capacity,), torch.float64)], device, group=group)
        self.region = region
        self.R, self.rank, self.cap = region.world_size, region.rank, int(capacity)
        self.off = 0

    def begin_step(self) -> None:
        self.off = 0

    def sync(self, tensors, scale: float = 1.0) -> None:
        """In place, for each contiguous fp64 tensor t (even element count): t <- scale * sum over ranks of t. ONE barrier for the lot."""
        R, taken = self.R, []
        for t in tensors:
            n = t.numel()
            if t.dtype != torch.float64 or not t.is_contiguous() or n % 2:
                raise ValueError("SyncStats.sync takes contiguous float64 tensors with an even element count")
            if self.off + n > self.cap:
                raise RuntimeError("SyncStats: slot capacity exhausted (was begin_step() called at the start of the step?)")
            base = self.off * R
            self.region.bcast(t, "bn", (base + self.rank * n) * 8)
            taken.append((base, n))
            self.off += n
        self.region.barrier(0)
        slots = self.region.local("bn")
        for t, (base, n) in zip(tensors, taken):
            L.call("pc_peer_sum_slots", L.ptr(slots[base:base + R * n], torch.float64), R, n, float(scale), L.ptr(t, torch.float64), L.stream())
