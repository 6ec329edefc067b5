"""Host logic of the row-sharded repeated SpMV (ellspmv_b200/sharded.py) on
CPU: world_size-2 and -3 gloo process groups, with the oracle standing in for
the per-shard kernel (the oracle is the checker AND the stand-in here because
there is no GPU; the product path never does this).  The sharded iteration
must reproduce the single-process oracle iteration bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT
from ellspmv_b200.sharded import exchanged_bytes, partition_rows, push_plan


def test_partition_rows_is_the_reference_static_split():
    assert partition_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]       # N/T + (p < N%T), csrspmv.c:2238
    assert partition_rows(8, 8) == [(i, i + 1) for i in range(8)]
    assert partition_rows(3, 5) == [(0, 1), (1, 2), (2, 3), (3, 3), (3, 3)]
    assert partition_rows(0, 2) == [(0, 0), (0, 0)]
    for n, w in [(452984832, 8), (67108864 * 4, 4), (1000003, 7)]:
        parts = partition_rows(n, w)
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
        assert max(b - a for a, b in parts) - min(b - a for a, b in parts) <= 1


def test_push_plan_restricts_to_what_peers_reference():
    parts = partition_rows(400, 4)
    # a stencil-like shard references its own rows +- 10
    needs = [(max(0, a - 10), min(400, b + 10)) for a, b in parts]
    assert push_plan(1, parts, needs) == [(0, 100, 110), (2, 190, 200)]
    assert push_plan(0, parts, needs) == [(1, 90, 100)]
    assert exchanged_bytes(1, parts, needs, "push") == 20 * 8
    assert exchanged_bytes(1, parts, needs, "allgather") == 100 * 8 * 3
    # a random matrix references everything: push degenerates to the all-gather volume
    needs = [(0, 400)] * 4
    assert exchanged_bytes(2, parts, needs, "push") == exchanged_bytes(2, parts, needs, "allgather")
    # an empty shard needs nothing and sends nothing
    assert push_plan(0, [(0, 0), (0, 5)], [(0, 0), (0, 5)]) == []


def test_sync_ranks_are_the_ranks_rows_are_exchanged_with():
    from ellspmv_b200.sharded import sync_ranks
    parts = partition_rows(400, 4)
    needs = [(max(0, a - 10), min(400, b + 10)) for a, b in parts]      # stencil: neighbours only
    assert [sync_ranks(r, parts, needs) for r in range(4)] == [[1], [0, 2], [1, 3], [2]]
    needs = [(0, 400)] * 4                                               # scattered: everybody
    assert sync_ranks(2, parts, needs) == [0, 1, 3]
    # one-directional: rank 1 reads rank 0's rows, rank 0 reads only its own -> both list each other
    # (rank 0 must not overwrite what rank 1 still reads; rank 1 must see rank 0's pushes)
    needs = [(0, 100), (50, 200)]
    parts2 = [(0, 100), (100, 200)]
    assert sync_ranks(0, parts2, needs) == [1] and sync_ranks(1, parts2, needs) == [0]
    # block diagonal: nobody to wait for
    assert sync_ranks(0, parts2, [(0, 100), (100, 200)]) == []


class _Info:
    pass


class OracleShard:
    """Duck-types EllMatrix for rows [lo, hi) with the oracle's ellgemv."""

    def __init__(self, oracle, K, ncols, ec, ea, lo, hi, global_rows):
        self.o, self.K, self.lo, self.hi = oracle, K, lo, hi
        self.ec = np.ascontiguousarray(ec[lo * K:hi * K])
        self.ea = np.ascontiguousarray(ea[lo * K:hi * K])
        self.i = _Info()
        self.i.global_rows, self.i.num_columns, self.i.row_begin, self.i.num_rows = global_rows, ncols, lo, hi - lo
        self.i.min_col = int(self.ec.min()) if len(self.ec) else 0
        self.i.max_col = int(self.ec.max()) if len(self.ec) else -1
        self.i.device = 0

    def info(self):
        return self.i

    def spmv_device(self, y, x, mode, stream=0):
        yn = np.zeros(self.hi - self.lo)
        self.o.ellgemv(self.hi - self.lo, yn, x.numpy(), self.K, self.ec, self.ea)
        y.copy_(torch.from_numpy(yn))


CASES = [("laplace2d", (12, 9), (0.25, -0.125), False),
         ("stencil27", (6, 5, 4), (0.5, -1.0 / 52), True),
         ("random", (91, 91, 7), (0, 0), True)]


def _worker(rank, world, port, iters, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from ellspmv_b200.sharded import ShardedIterate, partition_rows
    from oracle.pyoracle import Oracle
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        orc = Oracle()
        for kind, dims, vals, uneven in CASES:
            K, ncols, ec, ea, _ = orc.gen_ell(kind, dims, vals, seed=42, bits=32)
            n = len(ea) // K
            parts = partition_rows(n, world)
            if uneven:     # move the first cut: ragged blocks take the broadcast path
                parts = [(0, parts[0][1] - 3)] + [(parts[1][0] - 3, parts[1][1])] + parts[2:]
            lo, hi = parts[rank]
            shard = OracleShard(orc, K, ncols, ec, ea, lo, hi, n)
            it = ShardedIterate(shard, rank, world, exchange="allgather", device=torch.device("cpu"))
            x0 = np.random.default_rng(0).uniform(-1, 1, n)
            it.set_x(lambda a, b: torch.from_numpy(x0[a:b].copy()))
            for _ in range(iters):
                it.step()
            got = it.gather_result().numpy()
            want = orc.ell_iterate(n, x0, iters, K, ec, ea)
            ok = np.array_equal(got.view(np.uint64), want.view(np.uint64))
            full_ok = np.array_equal(it.current().numpy().view(np.uint64), want.view(np.uint64))
            q.put((rank, kind, bool(ok and full_ok), it.describe()["mode"], it.parts == parts))
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_sharded_iteration_equals_single_process_oracle():
    """world_size 2 over gloo; even and ragged row blocks; 5 iterations each."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 5, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(world * len(CASES))]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(results) == world * len(CASES)
    assert all(r[2] and r[3] == "allgather" and r[4] for r in results), results
