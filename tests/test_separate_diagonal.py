"""Separate-diagonal variants: ellgemvsd / ellgemv16sd
(ellspmv.c:1155-1221) and csrgemvsd (csrspmv.c:1598-1629).

Golden vectors come from the unmodified reference: its ELL functions called
with their flags in DECLARED order (its main() swaps them, Q1), its CSR
functions, and the csrspmv program's --separate-diagonal stdout."""
import os
import subprocess

import numpy as np
import pytest

import hostlib
from conftest import bits_equal, load_golden, unhex

SD_CASES = ["sd_rand", "sd_k16"]


def arrays(g, bits):
    dt = np.int32 if bits == 32 else np.int64
    return (np.array(g["rowidx"], dtype=dt), np.array(g["colidx"], dtype=dt), unhex(g["a"]), unhex(g["x"]),
            unhex(g["y0"]))


@pytest.mark.parametrize("name", SD_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_oracle_matches_reference(oracle, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    n = g["num_rows"]
    ri, ci, a, x, y0 = arrays(g, bits)
    K, ellsize, diagsize, ec, ea, ad = oracle.ell_from_coo_sd(n, n, ri, ci, a)
    assert (K, ellsize, diagsize) == (e["rowsize"], e["ellsize"], e["diagsize"])
    assert ec.tolist() == e["ellcolidx"] and bits_equal(ea, unhex(e["ella"])) and bits_equal(ad, unhex(e["ellad"]))
    y = y0.copy()
    oracle.ellgemvsd(n, y, x, K, ec, ea, ad, 0)
    assert bits_equal(y, unhex(e["y_ellgemvsd"]))
    if e["y_ellgemv16sd"] is not None:
        assert K == 16
        y = y0.copy()
        oracle.ellgemvsd(n, y, x, K, ec, ea, ad, 1)
        assert bits_equal(y, unhex(e["y_ellgemv16sd"]))
    rowptr, cc, ca, cad, lo, hi = oracle.csr_from_coo_sd(n, ri, ci, a)
    assert rowptr.tolist() == e["rowptr"] and cc.tolist() == e["csrcolidx"]
    assert bits_equal(ca, unhex(e["csra"])) and bits_equal(cad, unhex(e["csrad"]))
    assert (lo, hi) == (e["rowsizemin"], e["rowsizemax"]) and e["csr_diagsize"] == n
    y = y0.copy()
    oracle.csrgemvsd(n, y, x, rowptr, cc, ca, cad)
    assert bits_equal(y, unhex(e["y_csrgemvsd"]))


@pytest.mark.parametrize("name", SD_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_host_converters_match_reference(tmp_path, name, bits):
    import ctypes as C
    g = load_golden(name)
    e = g[f"idx{bits}"]
    n = g["num_rows"]
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, n, n, g["rowidx"], g["colidx"], unhex(g["a"]))
    lib = hostlib.hostlib(bits)
    it = C.c_int32 if bits == 32 else C.c_int64
    dt = np.int32 if bits == 32 else np.int64
    dims = (C.c_int64 * 7)()
    colidx, a, ad = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.host_ell_from_file_sd(p.encode(), 0, 1, dims, C.byref(colidx), C.byref(a), C.byref(ad)) == 0
    assert list(dims)[3:6] == [e["rowsize"], e["ellsize"], e["diagsize"]]
    assert hostlib._take(lib, colidx, dims[4], it, dt).tolist() == e["ellcolidx"]
    assert bits_equal(hostlib._take(lib, a, dims[4], C.c_double, np.float64), unhex(e["ella"]))
    assert bits_equal(hostlib._take(lib, ad, dims[5], C.c_double, np.float64), unhex(e["ellad"]))
    rowptr = C.c_void_p()
    assert lib.host_csr_from_file_sd(p.encode(), 0, 1, dims, C.byref(rowptr), C.byref(colidx), C.byref(a), C.byref(ad)) == 0
    assert list(dims)[3:6] == [len(e["csrcolidx"]), e["rowsizemin"], e["rowsizemax"]]
    assert hostlib._take(lib, rowptr, n + 1, C.c_int64, np.int64).tolist() == e["rowptr"]
    assert hostlib._take(lib, colidx, dims[3], it, dt).tolist() == e["csrcolidx"]
    assert bits_equal(hostlib._take(lib, a, dims[3], C.c_double, np.float64), unhex(e["csra"]))
    assert bits_equal(hostlib._take(lib, ad, n, C.c_double, np.float64), unhex(e["csrad"]))


# ------------------------------- GPU ----------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("name", SD_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_gpu_golden(lib, name, bits):
    import ellspmv_b200 as E
    g = load_golden(name)
    e = g[f"idx{bits}"]
    n = g["num_rows"]
    dt = np.int32 if bits == 32 else np.int64
    x, y0 = unhex(g["x"]), unhex(g["y0"])
    ec, ea, ad = np.array(e["ellcolidx"], dtype=dt), unhex(e["ella"]), unhex(e["ellad"])
    for R in (1, 2, 4):
        A = E.EllMatrix.upload(n, n, e["rowsize"], ec, ea, E.rows_per_thread(R))
        A.set_diagonal(ad, 0)
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        assert bits_equal(y, unhex(e["y_ellgemvsd"])), (name, bits, R)
        if e["y_ellgemv16sd"] is not None:
            A.set_diagonal(ad, 1)
            y = y0.copy()
            A.spmv(y, x, 1, E.ACCUMULATE)
            assert bits_equal(y, unhex(e["y_ellgemv16sd"]))
        A.set_diagonal(None)            # detached again: plain ellgemv on the off-diagonal part
        A.free()
    Cm = E.CsrMatrix.upload(n, n, np.array(e["rowptr"], dtype=np.int64), np.array(e["csrcolidx"], dtype=dt), unhex(e["csra"]))
    Cm.set_diagonal(unhex(e["csrad"]))
    y = y0.copy()
    Cm.spmv(y, x, 1, E.ACCUMULATE)
    Cm.free()
    assert bits_equal(y, unhex(e["y_csrgemvsd"]))


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("shape", [(3000, 3000, 200000), (700, 900, 9000), (1, 1, 1)])
def test_gpu_vs_oracle(lib, oracle, bits, shape, monkeypatch):
    import ellspmv_b200 as E
    monkeypatch.setenv("ELLSPMV_CUDA_BLOCK_BYTES", "2048")   # several column blocks for the staged-gather pass
    nr, nc, nnz = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + nnz + bits)
    ri = rng.integers(1, nr + 1, nnz).astype(dt)
    ci = rng.integers(1, nc + 1, nnz).astype(dt)
    ci[: nnz // 5] = np.minimum(ri[: nnz // 5], nc)        # many diagonal entries
    a = rng.standard_normal(nnz)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    K, _, _, ec, ea, ad = oracle.ell_from_coo_sd(nr, nc, ri, ci, a)
    for order in (0, 1):
        want = y0.copy()
        oracle.ellgemvsd(nr, want, x, K, ec, ea, ad, order)
        want0 = np.zeros(nr)
        oracle.ellgemvsd(nr, want0, x, K, ec, ea, ad, order)
        for flags in (0, E.STAGED_GATHER):
            A = E.EllMatrix.upload(nr, nc, K, ec, ea, flags)
            A.set_diagonal(ad, order)
            y = y0.copy()
            A.spmv(y, x, 1, E.ACCUMULATE)
            assert bits_equal(y, want), (order, flags)
            A.spmv(y, x, 1, E.OVERWRITE)
            assert bits_equal(y, want0), (order, flags)
            A.free()
    # tolerance modes
    absprod = np.abs(ea.reshape(nr, K) * x[ec.reshape(nr, K)]).sum(axis=1) + np.abs(ad[:nr] * x[:nr])
    for flags in (E.FMA, E.KERNEL_WARP):
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, flags)
        A.set_diagonal(ad, 0)
        y = np.zeros(nr)
        A.spmv(y, x, 1, E.ACCUMULATE)
        A.free()
        assert np.all(np.abs(y - want0) <= (K + 3) * 2.0 ** -53 * absprod + 1e-300)
    if nr == nc:
        rowptr, cc, ca, cad, _, _ = oracle.csr_from_coo_sd(nr, ri, ci, a)
        want = y0.copy()
        oracle.csrgemvsd(nr, want, x, rowptr, cc, ca, cad)
        Cm = E.CsrMatrix.upload(nr, nc, rowptr, cc, ca)
        Cm.set_diagonal(cad)
        y = y0.copy()
        Cm.spmv(y, x, 1, E.ACCUMULATE)
        Cm.free()
        assert bits_equal(y, want)


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [32, 64])
def test_gpu_separate_diagonal_on_several_gpus(lib, oracle, bits):
    """csrgemvsd / ellgemvsd over row shards: every shard multiplies its diagonal entry with x at
    the row's GLOBAL index (the CSR shards used the shard-local row: wrong for every shard but
    the first)."""
    import torch

    import ellspmv_b200 as E
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    n = min(torch.cuda.device_count(), 4)
    nr = nc = 3001
    nnz = 90000
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(77 + bits)
    ri = rng.integers(1, nr + 1, nnz).astype(dt)
    ci = rng.integers(1, nc + 1, nnz).astype(dt)
    ci[: nnz // 5] = ri[: nnz // 5]
    a = rng.standard_normal(nnz)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    rowptr, cc, ca, cad, _, _ = oracle.csr_from_coo_sd(nr, ri, ci, a)
    want = y0.copy()
    oracle.csrgemvsd(nr, want, x, rowptr, cc, ca, cad)
    Cm = E.CsrMatrix.upload(nr, nc, rowptr, cc, ca, num_gpus=n)
    Cm.set_diagonal(cad)
    y = y0.copy()
    Cm.spmv(y, x, 1, E.ACCUMULATE)
    Cm.free()
    assert bits_equal(y, want)
    K, _, _, ec, ea, ad = oracle.ell_from_coo_sd(nr, nc, ri, ci, a)
    for order in (0, 1):
        want = y0.copy()
        oracle.ellgemvsd(nr, want, x, K, ec, ea, ad, order)
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, num_gpus=n)
        A.set_diagonal(ad, order)
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        A.free()
        assert bits_equal(y, want), order


@pytest.mark.gpu
def test_gpu_rejects_more_rows_than_columns(lib):
    import ellspmv_b200 as E
    A = E.EllMatrix.upload(5, 3, 1, np.zeros(5, dtype=np.int32), np.ones(5))
    with pytest.raises(E.EllspmvCudaError):
        A.set_diagonal(np.ones(5), 0)
    A.free()


@pytest.mark.gpu
@pytest.mark.parametrize("name", SD_CASES)
def test_gpu_host_programs(tmp_path, oracle, name):
    hostlib.build_host()
    g = load_golden(name)
    n = g["num_rows"]
    A = str(tmp_path / "A.mtx")
    xf, yf = str(tmp_path / "x.mtx"), str(tmp_path / "y.mtx")
    hostlib.write_mtx(A, n, n, g["rowidx"], g["colidx"], unhex(g["a"]), comments=())
    hostlib.write_vec(xf, unhex(g["x"]))
    hostlib.write_vec(yf, unhex(g["y0"]))
    env = dict(os.environ, LC_ALL="C")

    def run(prog, args):
        return subprocess.run([os.path.join(hostlib.BIN, prog)] + args, capture_output=True, text=True, env=env)
    # CSR: byte-identical to the reference program's stdout
    r = run("csrspmv", ["--separate-diagonal", A])
    assert r.returncode == 0 and r.stdout == g["program"]["csrspmv_sd"]["stdout"]
    r = run("csrspmv64", ["--separate-diagonal", "--repeat=2", A, xf, yf])
    assert r.returncode == 0 and r.stdout == g["program"]["csrspmv64_sd_xy"]["stdout"]
    # ELL: the reference program is broken here (Q1); expect its declared-order functions' result
    e = g["idx32"]
    want = unhex(g["y0"])
    order = 1 if e["rowsize"] == 16 else 0
    oracle.ellgemvsd(n, want, unhex(g["x"]), e["rowsize"], np.array(e["ellcolidx"], dtype=np.int32), unhex(e["ella"]),
                     unhex(e["ellad"]), order)
    r = run("ellspmv", ["--separate-diagonal", "-v", A, xf, yf])
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines()[2:] == ["%.15g" % v for v in want]
    assert ("gemv16sd: " if order else "gemvsd: ") in r.stderr
