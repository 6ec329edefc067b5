"""Row-sharded repeated SpMV on real GPUs (needs >= 2): every exchange mode
gives the single-GPU bits.  One process per GPU over NCCL, like bench.py."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

import ellspmv_b200 as E
from conftest import ROOT, bits_equal

pytestmark = pytest.mark.gpu

CASES = [(E.GEN_LAPLACE2D, "laplace2d", (64, 100), (0.25, -0.125), 32),
         # many slices per shard: interior CTAs (which never wait in the fused step hand-shake)
         # next to halo CTAs, and a row split that is not a multiple of 16 entries
         (E.GEN_LAPLACE2D, "laplace2d", (1201, 97), (0.25, -0.125), 32),
         (E.GEN_STENCIL27, "stencil27", (24, 9, 11), (0.5, -1.0 / 52), 64),
         (E.GEN_RANDOM, "random", (5003, 5003, 9), (0.0, 0.0), 32)]
MODES = ("allgather", "push-neighbours", "push-fusedsync", "push-device", "push-nccl")
STEPS = 12


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from ellspmv_b200.sharded import ShardedIterate, partition_rows
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        for kind, name, dims, vals, bits in CASES:
            rows = dims[0] if name == "random" else int(np.prod(dims))
            x0 = np.random.default_rng(0).uniform(-1, 1, rows)
            for mode in MODES:
                lo, hi = partition_rows(rows, world)[rank]
                fused = mode.endswith("fusedsync")        # the hand-shake inside the SpMV kernel (opt-in upload flag)
                A = E.EllMatrix.generate(kind, dims, vals, 42, bits, row_begin=lo, row_end=hi, device=rank,
                                         flags=E.FUSED_SYNC if fused else 0)
                it = ShardedIterate(A, rank, world, exchange=mode.split("-")[0],
                                    barrier="neighbours" if fused or "-" not in mode else mode.split("-")[1])
                it.set_x(lambda a, b: torch.from_numpy(x0[a:b].copy()).to(dev))
                for _ in range(STEPS):
                    it.step(torch.cuda.current_stream().cuda_stream)
                got = it.gather_result().cpu().numpy()
                q.put((rank, name + str(dims), mode, got if rank == 0 else None, it.describe()))
                it.close()
                A.free()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_sharded_equals_single_gpu(lib, oracle):
    world = min(torch.cuda.device_count(), 4)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=600) for _ in range(world * len(CASES) * len(MODES))]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    names = {"laplace2d": "laplace2d", "stencil27": "stencil27", "random": "random"}
    for kind, name, dims, vals, bits in CASES:
        K, ncols, ec, ea, _ = oracle.gen_ell(names[name], dims, vals, seed=42, bits=bits)
        rows = len(ea) // K
        x0 = np.random.default_rng(0).uniform(-1, 1, rows)
        want = oracle.ell_iterate(rows, x0, STEPS, K, ec, ea)
        for mode in MODES:
            got = [r for r in results if r[0] == 0 and r[1] == name + str(dims) and r[2] == mode]
            assert len(got) == 1 and bits_equal(got[0][3], want), (name, dims, mode)
            d = got[0][4]
            if name != "random" and mode.startswith("push"):
                # a stencil shard references only a halo: far less than the all-gather volume
                assert d["bytes_sent_per_step_rank0"] < rows * 8 // world
