"""The CPU oracle against the golden vectors recorded from the unmodified
reference (tests/golden/*.json, made by tests/golden/make_golden.py).
Bit-exact: integer arrays equal, fp64 arrays equal as 64-bit patterns."""
import numpy as np
import pytest

from conftest import GOLDEN_CASES, bits_equal, load_golden, unhex


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_ell_from_coo_matches_reference(oracle, name, bits):
    g = load_golden(name)
    dt = np.int32 if bits == 32 else np.int64
    ri, ci, a = np.array(g["rowidx"], dtype=dt), np.array(g["colidx"], dtype=dt), unhex(g["a"])
    K, ellsize, diagsize, ec, ea = oracle.ell_from_coo(g["num_rows"], g["num_columns"], ri, ci, a)
    e = g[f"idx{bits}"]
    assert (K, ellsize, diagsize) == (e["rowsize"], e["ellsize"], e["diagsize"])
    assert np.array_equal(ec, np.array(e["ellcolidx"], dtype=dt))
    assert bits_equal(ea, unhex(e["ella"]))


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_ellgemv_matches_reference(oracle, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    dt = np.int32 if bits == 32 else np.int64
    ec, ea = np.array(e["ellcolidx"], dtype=dt), unhex(e["ella"])
    x = unhex(g["x"])
    y = unhex(g["y0"])
    oracle.ellgemv(g["num_rows"], y, x, e["rowsize"], ec, ea)
    assert bits_equal(y, unhex(e["y_ell"]))
    oracle.ellgemv(g["num_rows"], y, x, e["rowsize"], ec, ea)
    oracle.ellgemv(g["num_rows"], y, x, e["rowsize"], ec, ea)
    assert bits_equal(y, unhex(e["y_ell_repeat3"]))   # y accumulates across repeats (Q6)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_csr_matches_reference(oracle, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    dt = np.int32 if bits == 32 else np.int64
    ri, ci, a = np.array(g["rowidx"], dtype=dt), np.array(g["colidx"], dtype=dt), unhex(g["a"])
    rowptr, cc, ca, lo, hi = oracle.csr_from_coo(g["num_rows"], g["num_columns"], ri, ci, a)
    assert np.array_equal(rowptr, np.array(e["rowptr"], dtype=np.int64))
    assert np.array_equal(cc, np.array(e["csrcolidx"], dtype=dt))
    assert bits_equal(ca, unhex(e["csra"]))
    assert (lo, hi) == (e["rowsizemin"], e["rowsizemax"])
    y = unhex(g["y0"])
    oracle.csrgemv(g["num_rows"], y, unhex(g["x"]), rowptr, cc, ca)
    assert bits_equal(y, unhex(e["y_csr"]))


def test_test_mtx_known_answers(oracle):
    """The values SURVEY.md 4 lists for the reference's only fixture."""
    g = load_golden("test_mtx")
    e = g["idx32"]
    assert (e["rowsize"], e["ellsize"], e["diagsize"]) == (5, 20, 4)
    assert e["ellcolidx"] == [1, 0, 0, 0, 0, 1, 1, 1, 1, 1, 2, 2, 2, 2, 2, 4, 0, 3, 2, 1]
    assert unhex(e["ella"]).tolist() == [2, 1, 0, 0, 0, 1, 0, 0, 0, 0, 3, 0, 0, 0, 0, 1, -1, 2, 3, 1]
    assert unhex(e["y_ell"]).tolist() == [3, 1, 3, 6]
    assert e["rowptr"] == [0, 2, 3, 4, 9]
    assert e["csrcolidx"] == [1, 0, 1, 2, 4, 0, 3, 2, 1]
    assert g["program"]["ellspmv"]["stdout"] == "%%MatrixMarket vector array real general\n4\n3\n1\n3\n6\n"
    assert g["program"]["csrspmv"]["stdout"] == g["program"]["ellspmv"]["stdout"]


def test_iterate_is_repeated_overwrite(oracle):
    rng = np.random.default_rng(3)
    n, K = 40, 4
    ec = rng.integers(0, n, n * K).astype(np.int32)
    ea = rng.uniform(-0.25, 0.25, n * K)
    x = rng.uniform(-1, 1, n)
    want = x.copy()
    for _ in range(3):
        y = np.zeros(n)
        oracle.ellgemv(n, y, want, K, ec, ea)
        want = y
    got = oracle.ell_iterate(n, x, 3, K, ec, ea)
    assert bits_equal(got, want)


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_rowsort_matches_reference(oracle, name, bits):
    """--sort-rows for CSR (csrspmv.c:1269-1388), including the tie order of duplicate columns."""
    g = load_golden(name)
    e = g[f"idx{bits}"]
    dt = np.int32 if bits == 32 else np.int64
    rowptr = np.array(e["rowptr"], dtype=np.int64)
    cc, ca = np.array(e["csrcolidx"], dtype=dt), unhex(e["csra"])
    oracle.rowsort(g["num_rows"], rowptr, cc, ca)
    assert cc.tolist() == e["csrcolidx_sorted"] and bits_equal(ca, unhex(e["csra_sorted"]))
    y = unhex(g["y0"])
    oracle.csrgemv(g["num_rows"], y, unhex(g["x"]), rowptr, cc, ca)
    assert bits_equal(y, unhex(e["y_csr_sorted"]))
