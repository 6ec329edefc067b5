"""One handle over several GPUs of one process (ellspmv_cuda_upload with
num_gpus > 1, the C host program's --gpus=N): the reference semantics
(y += A*x, x constant) and the y -> x iteration give the single-GPU bits.
Needs >= 2 GPUs."""
import os
import subprocess

import numpy as np
import pytest
import torch

import ellspmv_b200 as E
import hostlib
from conftest import bits_equal, load_golden, unhex

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")]


def ngpus():
    return min(torch.cuda.device_count(), 4)


@pytest.mark.parametrize("kind,name,dims,vals,bits", [
    (E.GEN_LAPLACE2D, "laplace2d", (61, 100), (0.5, 0.125), 32),
    (E.GEN_STENCIL27, "stencil27", (23, 9, 11), (0.5, 1.0 / 52), 64),
    (E.GEN_RANDOM, "random", (5003, 5003, 9), (0.0, 0.0), 32)])
@pytest.mark.parametrize("flags", [0, E.STAGED_GATHER, E.FUSED_SYNC])
def test_group_equals_oracle(lib, oracle, kind, name, dims, vals, bits, flags, monkeypatch):
    # STAGED_GATHER: every shard gathers x column block by column block (several blocks at this
    # size with 4 KB per block) and the fused push runs out of the staged sum kernel
    monkeypatch.setenv("ELLSPMV_CUDA_BLOCK_BYTES", "4096")
    n = ngpus()
    K, ncols, ec, ea, _ = oracle.gen_ell(name, dims, vals, seed=42, bits=bits)
    rows = len(ea) // K
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, ncols)
    y0 = rng.uniform(-1, 1, rows)
    want = y0.copy()
    for _ in range(3):
        oracle.ellgemv(rows, want, x, K, ec, ea)
    for A in (E.EllMatrix.upload(rows, ncols, K, ec, ea, flags, num_gpus=n),
              E.EllMatrix.generate(kind, dims, vals, 42, bits, flags=flags, num_gpus=n)):
        i = A.info()
        assert (i.num_gpus, i.num_rows, i.rowsize) == (n, rows, K)
        assert (i.min_col, i.max_col) == (ec.min(), ec.max())
        c2, a2 = A.download()
        assert np.array_equal(c2, ec) and bits_equal(a2, ea)
        y = y0.copy()
        secs = A.spmv(y, x, 3, E.ACCUMULATE)
        assert bits_equal(y, want) and np.all(secs > 0)
        y = np.full(rows, 7.0)
        A.spmv(y, x, 1, E.OVERWRITE)
        w0 = np.zeros(rows)
        oracle.ellgemv(rows, w0, x, K, ec, ea)
        assert bits_equal(y, w0)
        for iters in (1, 6):
            y = np.zeros(rows)
            A.spmv(y, x, iters, E.ITERATE)
            assert bits_equal(y, oracle.ell_iterate(rows, x, iters, K, ec, ea)), (name, iters)
        A.free()


def test_group_rejects_device_vector_calls(lib):
    A = E.EllMatrix.generate(E.GEN_LAPLACE2D, (40, 40), (4.0, -1.0), 42, 32, num_gpus=2)
    x = torch.ones(1600, dtype=torch.float64, device="cuda")
    with pytest.raises(E.EllspmvCudaError):
        A.spmv_device(x, x, E.OVERWRITE, 0)
    A.free()


def test_host_program_gpus_option(tmp_path, oracle):
    hostlib.build_host()
    n = ngpus()
    env = dict(os.environ, LC_ALL="C")
    g = load_golden("rand_square")
    A = str(tmp_path / "A.mtx")
    hostlib.write_mtx(A, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]), comments=())
    r = subprocess.run([os.path.join(hostlib.BIN, "ellspmv"), f"--gpus={n}", "--repeat=2", "--warmup=1", "-v", A],
                       capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout == g["program"]["ellspmv_repeat2_warmup1"]["stdout"]
    assert f"{n} GPU(s)" in r.stderr
    K, ncols, ec, ea, _ = oracle.gen_ell("stencil27", (16, 7, 5), (0.5, -1.0 / 52), bits=64)
    want = oracle.ell_iterate(ncols, np.ones(ncols), 5, K, ec, ea)
    r = subprocess.run([os.path.join(hostlib.BIN, "ellspmv64"), f"--gpus={n}", "--synthetic=stencil27s:16,7,5",
                        "--iterate", "--repeat=5"], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines()[2:] == ["%.15g" % v for v in want]


def test_csr_group(lib, oracle, tmp_path):
    """CSR over several GPUs: nonzero-balanced contiguous row blocks (csrspmv.c:1700-1708)."""
    n = ngpus()
    rng = np.random.default_rng(8)
    nr, nc = 20011, 15000
    lens = rng.integers(0, 30, nr)
    lens[:50] = 4000                                    # a heavy head: the balanced cut is far from nr/n
    rowptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    for dt in (np.int32, np.int64):
        cc = rng.integers(0, nc, nnz).astype(dt)
        ca = rng.standard_normal(nnz)
        x = rng.standard_normal(nc)
        y0 = rng.standard_normal(nr)
        want = y0.copy()
        for _ in range(2):
            oracle.csrgemv(nr, want, x, rowptr, cc, ca)
        for flags in (0, E.KERNEL_THREAD, 3):
            A = E.CsrMatrix.upload(nr, nc, rowptr, cc, ca, flags, num_gpus=n)
            y = y0.copy()
            secs = A.spmv(y, x, 2, E.ACCUMULATE)
            assert bits_equal(y, want) and np.all(secs > 0)
            A.free()
    hostlib.build_host()
    g = load_golden("rand_square")
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]), comments=())
    r = subprocess.run([os.path.join(hostlib.BIN, "csrspmv"), f"--gpus={n}", p], capture_output=True, text=True,
                       env=dict(os.environ, LC_ALL="C"))
    assert r.returncode == 0 and r.stdout == g["program"]["csrspmv"]["stdout"]
