"""Row-sharded repeated SpMV, one process per GPU (BASELINE config 5).

    x_{k+1} <- A * x_k,   rows of A in contiguous blocks, one block per rank

Every rank keeps two full-length vectors (current x, next x).  A step runs the
local SpMV over the rank's rows, writing y straight into its slice of the next
x, and makes the slices other ranks need visible to them:

  exchange="allgather"  NCCL all-gather of the slices over NVLink
                        (torch.distributed plumbing), the north-star baseline;
  exchange="push"       the SpMV kernel itself stores each fresh y[i] into the
                        peers' next-x vectors through peer-mapped memory
                        (CUDA IPC + NVLink), restricted to the row range each
                        peer's shard actually references (its [min_col,
                        max_col]); a one-element all-reduce orders the steps.
                        For a stencil this moves only the halo planes, for a
                        random matrix it degenerates to the full all-gather.

Row sharding leaves every row's summation order untouched, so the result is
bit-identical to the single-GPU one (tested).

The reference has no multi-process path; the partition rule is its static
OpenMP split, rows/T + (p < rows % T) (csrspmv.c:2238).  This module is host
logic only -- the compute is `EllMatrix.spmv_device` / `spmv_push`.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from . import OVERWRITE, EllspmvCudaError, load_library

Range = Tuple[int, int]


def partition_rows(global_rows: int, world: int) -> List[Range]:
    """Contiguous row blocks, sizes rows//world + (p < rows % world)."""
    base, rem = divmod(global_rows, world)
    out, lo = [], 0
    for p in range(world):
        hi = lo + base + (1 if p < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def push_plan(rank: int, parts: Sequence[Range], needs: Sequence[Range]) -> List[Tuple[int, int, int]]:
    """(peer, lo, hi): the part of this rank's rows that peer's shard references."""
    my_lo, my_hi = parts[rank]
    plan = []
    for p, (nlo, nhi) in enumerate(needs):
        if p == rank:
            continue
        lo, hi = max(my_lo, nlo), min(my_hi, nhi)
        if lo < hi:
            plan.append((p, lo, hi))
    return plan


def sync_ranks(rank: int, parts: Sequence[Range], needs: Sequence[Range]) -> List[int]:
    """The ranks this one exchanges rows with: those it pushes to and those that push to it.
    These are the ranks whose previous step must be over before this rank's halo CTAs start
    (their pushes have landed; they no longer read the vector this step overwrites)."""
    out = {p for p, _, _ in push_plan(rank, parts, needs)}
    for p in range(len(parts)):
        if p != rank and any(q == rank for q, _, _ in push_plan(p, parts, needs)):
            out.add(p)
    return sorted(out)


def exchanged_bytes(rank: int, parts: Sequence[Range], needs: Sequence[Range], mode: str) -> int:
    """Bytes this rank sends per step."""
    if mode == "allgather":
        return (parts[rank][1] - parts[rank][0]) * 8 * (len(parts) - 1)
    return sum((hi - lo) * 8 for _, lo, hi in push_plan(rank, parts, needs))


class _DeviceBuffer:
    """A cudaMalloc'ed vector owned by the library (exportable over CUDA IPC,
    unlike a sub-allocated torch tensor), viewed as a torch tensor."""

    def __init__(self, n: int, device: torch.device):
        import ctypes as C
        lib = load_library()
        p = C.c_void_p()
        err = lib.ellspmv_cuda_malloc_device(C.byref(p), max(n, 1) * 8)
        if err:
            raise EllspmvCudaError(err, "ellspmv_cuda_malloc_device", lib.ellspmv_cuda_last_error().decode())
        self.ptr, self.n = p.value, n
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (self.ptr, False),
                                         "version": 3, "strides": None}
        self.tensor = torch.as_tensor(self, device=device)

    def ipc_handle(self) -> bytes:
        import ctypes as C
        buf = (C.c_ubyte * 64)()
        lib = load_library()
        err = lib.ellspmv_cuda_ipc_export(self.ptr, buf)
        if err:
            raise EllspmvCudaError(err, "ellspmv_cuda_ipc_export", lib.ellspmv_cuda_last_error().decode())
        return bytes(buf)

    def free(self):
        if self.ptr:
            self.tensor = None
            load_library().ellspmv_cuda_free_device(self.ptr)
            self.ptr = 0


def _ipc_open(handle: bytes) -> int:
    import ctypes as C
    lib = load_library()
    buf = (C.c_ubyte * 64).from_buffer_copy(handle)
    p = C.c_void_p()
    err = lib.ellspmv_cuda_ipc_open(buf, C.byref(p))
    if err:
        raise EllspmvCudaError(err, "ellspmv_cuda_ipc_open", lib.ellspmv_cuda_last_error().decode())
    return p.value


class ShardedIterate:
    """Repeated x <- A*x over row shards.  `A` is this rank's shard: an
    EllMatrix (or any object with .info(), .spmv_device(), .spmv_push())."""

    def __init__(self, A, rank: int, world: int, exchange: str = "auto", group=None,
                 device: Optional[torch.device] = None, barrier: str = "neighbours"):
        self.A, self.rank, self.world, self.group = A, rank, world, group
        # "neighbours": ellspmv_cuda_spmv_exchange -- the step hand-shake only with the ranks rows are exchanged
        #   with (a one-warp kernel after the SpMV kernel; inside it for handles uploaded with FUSED_SYNC);
        # "device": all-ranks flag barrier kernel after the kernel; "nccl": 1-element all-reduce
        if barrier == "fused":
            barrier = "neighbours"
        if barrier not in ("neighbours", "device", "nccl"):
            raise ValueError(f"unknown barrier {barrier!r}")
        self.barrier = barrier
        info = A.info()
        self.global_rows = int(info.global_rows)
        if int(info.num_columns) != self.global_rows:
            raise ValueError("iterating x <- A*x needs a square matrix")
        mine = (int(info.row_begin), int(info.row_begin) + int(info.num_rows))
        need = (int(info.min_col), int(info.max_col) + 1) if info.max_col >= info.min_col else (0, 0)
        gathered: List = [None] * world
        if world > 1:
            dist.all_gather_object(gathered, (mine, need), group=group)
        else:
            gathered = [(mine, need)]
        self.parts: List[Range] = [g[0] for g in gathered]
        self.needs: List[Range] = [g[1] for g in gathered]
        lo = 0
        for p, (a, b) in enumerate(self.parts):      # shards must tile [0, rows) in rank order
            if a != lo or b < a:
                raise ValueError(f"rank {p} holds rows [{a}, {b}); expected a block starting at {lo}")
            lo = b
        if lo != self.global_rows:
            raise ValueError("row shards do not cover the matrix")
        self.lo, self.hi = mine
        self.is_cuda = device is None or device.type == "cuda"
        self.device = device if device is not None else torch.device("cuda", int(info.device))
        if exchange == "auto":
            exchange = "push" if (self.is_cuda and world > 1) else "allgather"
        if exchange == "push" and not self.is_cuda:
            raise ValueError("exchange='push' needs CUDA peer memory")
        self.exchange = exchange
        self.equal_parts = len({b - a for a, b in self.parts}) == 1
        n = self.global_rows
        if self.is_cuda:
            self._bufs = [_DeviceBuffer(n, self.device), _DeviceBuffer(n, self.device)]
            self.x = [b.tensor for b in self._bufs]
        else:
            self._bufs = []
            self.x = [torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)]
        self.cur = 0
        self.steps_done = 0
        self.plan = push_plan(rank, self.parts, self.needs)
        self._peer_ptrs: List[List[int]] = [[], []]
        self._opened: List[int] = []
        if exchange == "push" and world > 1:
            handles: List = [None] * world
            dist.all_gather_object(handles, [b.ipc_handle() for b in self._bufs], group=group)
            opened = {}
            for p, _, _ in self.plan:
                opened[p] = [_ipc_open(h) for h in handles[p]]
                self._opened.extend(opened[p])
            for buf in (0, 1):
                self._peer_ptrs[buf] = [opened[p][buf] for p, _, _ in self.plan]
            self._flag = torch.zeros(1, dtype=torch.int32, device=self.device)
            # step barrier on the device: 32 int64 flags per rank, every rank maps all of them
            self._flags = _DeviceBuffer(32, self.device)
            self._flags.tensor.zero_()
            torch.cuda.synchronize(self.device)
            fh: List = [None] * world
            dist.all_gather_object(fh, self._flags.ipc_handle(), group=group)
            self._flag_ptrs = []
            for p in range(world):
                if p == rank:
                    self._flag_ptrs.append(self._flags.ptr)
                else:
                    ptr = _ipc_open(fh[p])
                    self._opened.append(ptr)
                    self._flag_ptrs.append(ptr)
            self._sync_ranks = sync_ranks(rank, self.parts, self.needs)
            dist.barrier(group=group)

    # -- data ---------------------------------------------------------------
    def set_x(self, fill: Callable[[int, int], torch.Tensor]) -> None:
        """Initialise the current x on every rank: fill(lo, hi) -> values of x[lo:hi]."""
        self.x[self.cur][:] = fill(0, self.global_rows)
        self.x[1 - self.cur].zero_()
        if self.is_cuda:
            torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def current(self) -> torch.Tensor:
        """The full current vector (valid on the ranges this rank needs, and
        everywhere after an all-gather exchange)."""
        return self.x[self.cur]

    def local(self) -> torch.Tensor:
        return self.x[self.cur][self.lo:self.hi]

    # -- one iteration ----------------------------------------------------------
    def step(self, stream: int = 0) -> None:
        cur, nxt = self.x[self.cur], self.x[1 - self.cur]
        y = nxt[self.lo:self.hi]
        if self.exchange == "push" and self.world > 1 and self.barrier == "neighbours":
            # SpMV + push, and the step hand-shake with the neighbouring ranks only
            self.A.spmv_exchange(y, cur, OVERWRITE, self._peer_ptrs[1 - self.cur],
                                 [lo for _, lo, _ in self.plan], [hi for _, _, hi in self.plan],
                                 self.rank, self._sync_ranks, [self._flag_ptrs[p] for p in self._sync_ranks],
                                 self._flags.ptr, self.steps_done + 1, stream)
        elif self.exchange == "push" and self.world > 1:
            self.A.spmv_push(y, cur, OVERWRITE, self._peer_ptrs[1 - self.cur],
                             [lo for _, lo, _ in self.plan], [hi for _, _, hi in self.plan], stream)
            # orders step k's pushes before step k+1's gathers on every rank
            if self.barrier == "device":
                self._peer_barrier(stream)
            else:
                dist.all_reduce(self._flag, group=self.group)
        else:
            self.A.spmv_device(y, cur, OVERWRITE, stream)
            if self.world > 1:
                self._allgather(nxt)
        self.cur = 1 - self.cur
        self.steps_done += 1

    def _peer_barrier(self, stream: int) -> None:
        import ctypes as C
        lib = load_library()
        ptrs = (C.c_void_p * self.world)(*self._flag_ptrs)
        err = lib.ellspmv_cuda_peer_barrier(self.rank, self.world, self.steps_done + 1, self._flags.ptr, ptrs,
                                            stream or None)
        if err:
            raise EllspmvCudaError(err, "ellspmv_cuda_peer_barrier", lib.ellspmv_cuda_last_error().decode())

    def _allgather(self, full: torch.Tensor) -> None:
        if self.equal_parts:
            dist.all_gather_into_tensor(full, full[self.lo:self.hi], group=self.group)   # in place
        else:
            for p, (a, b) in enumerate(self.parts):                                     # ragged split
                if b > a:
                    src = p if self.group is None else dist.get_global_rank(self.group, p)
                    dist.broadcast(full[a:b], src=src, group=self.group)

    def check(self) -> None:
        """Raise if a step synchronisation on the device gave up waiting for a peer
        (slot 16 of the flag array, written after ~20 s): the vectors are then stale."""
        if getattr(self, "_flags", None) is None:
            return
        if self.is_cuda:
            torch.cuda.synchronize(self.device)
        mark = int(self._flags.tensor.view(torch.int64)[16].item()) & 0xFFFFFFFF
        if mark:
            raise EllspmvCudaError(5, "ShardedIterate",   # EIO
                                   f"rank {self.rank} gave up waiting for rank {mark - 1} in the step synchronisation")

    def gather_result(self) -> torch.Tensor:
        """Full current vector assembled on every rank (for checking)."""
        self.check()
        out = self.x[self.cur].clone()
        if self.world > 1:
            self._allgather(out)
        return out

    def describe(self) -> dict:
        sent = exchanged_bytes(self.rank, self.parts, self.needs, self.exchange)
        return {"mode": self.exchange, "barrier": self.barrier if self.exchange == "push" else "nccl collective",
                "bytes_sent_per_step_rank0": sent,
                "rows_pushed_to_peers": [(p, hi - lo) for p, lo, hi in self.plan] if self.exchange == "push" else None,
                "needs": self.needs[self.rank]}

    def close(self) -> None:
        lib = load_library()
        try:
            self.check()
        finally:
            self._release(lib)

    def _release(self, lib) -> None:
        for p in self._opened:
            lib.ellspmv_cuda_ipc_close(p)
        self._opened = []
        self.x = []
        for b in self._bufs:
            b.free()
        self._bufs = []
        if getattr(self, "_flags", None) is not None:
            self._flags.free()
            self._flags = None
