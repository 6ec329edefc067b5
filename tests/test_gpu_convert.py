"""COO -> ELL / CSR on the device: bit-identical to
the reference's serial converters (golden vectors) and to the oracle on
random input with duplicates, empty rows and non-square shapes."""
import os
import subprocess

import numpy as np
import pytest

import ellspmv_b200 as E
import hostlib
from conftest import GOLDEN_CASES, bits_equal, load_golden, unhex

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_golden(lib, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    dt = np.int32 if bits == 32 else np.int64
    ri, ci, a = np.array(g["rowidx"], dtype=dt), np.array(g["colidx"], dtype=dt), unhex(g["a"])
    A = E.EllMatrix.upload_coo(g["num_rows"], g["num_columns"], ri, ci, a)
    i = A.info()
    assert i.rowsize == e["rowsize"]
    ec, ea = A.download()
    assert ec.tolist() == e["ellcolidx"] and bits_equal(ea, unhex(e["ella"]))
    y = unhex(g["y0"])
    A.spmv(y, unhex(g["x"]), 1, E.ACCUMULATE)
    assert bits_equal(y, unhex(e["y_ell"]))
    A.free()
    Cm = E.CsrMatrix.upload_coo(g["num_rows"], g["num_columns"], ri, ci, a)
    rowptr, cc, ca = Cm.download(g["num_rows"], len(a), bits)
    assert rowptr.tolist() == e["rowptr"] and cc.tolist() == e["csrcolidx"] and bits_equal(ca, unhex(e["csra"]))
    y = unhex(g["y0"])
    Cm.spmv(y, unhex(g["x"]), 1, E.ACCUMULATE)
    assert bits_equal(y, unhex(e["y_csr"]))
    Cm.free()


@pytest.mark.parametrize("shape", [(1000, 1000, 20000), (5000, 300, 60000), (300, 5000, 60000), (100000, 100000, 700000),
                                   (7, 3, 0)])
@pytest.mark.parametrize("bits", [32, 64])
def test_random_vs_oracle(lib, oracle, shape, bits):
    nr, nc, nnz = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + nc + nnz + bits)
    ri = rng.integers(1, nr + 1, nnz).astype(dt)
    ci = rng.integers(1, nc + 1, nnz).astype(dt)
    if nnz > 100:
        ri[50:60] = ri[40]; ci[50:60] = ci[40]              # duplicates of one entry
        ri[ri % 7 == 0] -= 1                                  # rows 7, 14, ... stay empty
        ri[:300] = 1                                          # one long row (K ~ 300 keeps N*K small)
    a = rng.standard_normal(nnz)
    assert np.bincount(ri, minlength=nr + 1).max() * nr < 80_000_000     # guard: the ELL arrays must stay small
    K, ellsize, _, ec, ea = oracle.ell_from_coo(nr, nc, ri, ci, a)
    for flags in (0, E.rows_per_thread(4), E.WIDE_INDEX):
        A = E.EllMatrix.upload_coo(nr, nc, ri, ci, a, flags)
        assert A.info().rowsize == K
        ec2, ea2 = A.download()
        assert np.array_equal(ec, ec2) and bits_equal(ea, ea2)
        if ellsize:
            assert (A.info().min_col, A.info().max_col) == (ec.min(), ec.max())
        A.free()
    rowptr, cc, ca, _, _ = oracle.csr_from_coo(nr, nc, ri, ci, a)
    Cm = E.CsrMatrix.upload_coo(nr, nc, ri, ci, a)
    rp2, cc2, ca2 = Cm.download(nr, nnz, bits)
    Cm.free()
    assert np.array_equal(rowptr, rp2) and np.array_equal(cc, cc2) and bits_equal(ca, ca2)


def test_out_of_range_indices_are_rejected(lib):
    ri = np.array([1, 4], dtype=np.int32)
    ci = np.array([1, 1], dtype=np.int32)
    with pytest.raises(E.EllspmvCudaError):
        E.EllMatrix.upload_coo(3, 3, ri, ci, np.ones(2))
    with pytest.raises(E.EllspmvCudaError):
        E.EllMatrix.upload_coo(4, 3, ri, np.array([1, 0], dtype=np.int32), np.ones(2))
    with pytest.raises(E.EllspmvCudaError):
        E.CsrMatrix.upload_coo(4, 3, ri, np.array([4, 1], dtype=np.int32), np.ones(2))


@pytest.mark.parametrize("name", ["test_mtx", "rand_wide", "rand_square"])
def test_host_programs_device_convert(tmp_path, name):
    hostlib.build_host()
    g = load_golden(name)
    A = str(tmp_path / "A.mtx")
    hostlib.write_mtx(A, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]), comments=())
    env = dict(os.environ, LC_ALL="C")
    for prog, key in (("ellspmv", "ellspmv"), ("ellspmv64", "ellspmv64"), ("csrspmv", "csrspmv"), ("csrspmv64", "csrspmv64")):
        r = subprocess.run([os.path.join(hostlib.BIN, prog), "--device-convert", "-v", A], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr
        assert r.stdout == g["program"][key]["stdout"]
        assert ("ell_from_coo: " if prog.startswith("ell") else "csr_from_coo: ") in r.stderr
