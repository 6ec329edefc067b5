"""Host side of the drop-in (no GPU needed): the Matrix Market reader and the
COO->ELL / COO->CSR converters of the host programs reproduce the reference's
arrays bit for bit (golden vectors), and the reader is as strict as the
reference's (ellspmv.c:707-888)."""
import errno
import gzip
import os
import subprocess

import numpy as np
import pytest

import hostlib
from conftest import GOLDEN_CASES, bits_equal, load_golden, unhex


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_host_ell_arrays_match_reference(tmp_path, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]))
    err, dims, ec, ea = hostlib.ell_from_file(bits, p)
    assert err == 0
    assert dims[:6] == [g["num_rows"], g["num_columns"], len(g["a"]), e["rowsize"], e["ellsize"], e["diagsize"]]
    assert ec.tolist() == e["ellcolidx"] and bits_equal(ea, unhex(e["ella"]))


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_host_csr_arrays_match_reference(tmp_path, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]))
    err, dims, rowptr, cc, ca = hostlib.csr_from_file(bits, p)
    assert err == 0
    assert dims[3:6] == [len(e["csrcolidx"]), e["rowsizemin"], e["rowsizemax"]]
    assert rowptr.tolist() == e["rowptr"] and cc.tolist() == e["csrcolidx"] and bits_equal(ca, unhex(e["csra"]))


def test_symmetric_expansion_matches_reference_csr(tmp_path):
    """csrspmv expands square symmetric input (csrspmv.c:1420-1427); checked
    against the unmodified reference when oracle/_ref is there."""
    from oracle.pyoracle import Reference
    if not Reference.available("csr", 32):
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(9)
    n, nnz = 30, 120
    ri = rng.integers(1, n + 1, nnz)
    ci = rng.integers(1, n + 1, nnz)
    lo = ri >= ci
    ri, ci = ri[lo], ci[lo]
    a = rng.standard_normal(len(ri))
    p = str(tmp_path / "S.mtx")
    hostlib.write_mtx(p, n, n, ri.tolist(), ci.tolist(), a, symmetry="symmetric")
    for bits in (32, 64):
        dt = np.int32 if bits == 32 else np.int64
        rowptr, cc, ca, lo_, hi_ = Reference("csr", bits).csr_from_coo(n, n, ri.astype(dt), ci.astype(dt), a, symmetric=True)
        err, dims, rp2, cc2, ca2 = hostlib.csr_from_file(bits, p)
        assert err == 0 and dims[4:6] == [lo_, hi_]
        assert np.array_equal(rowptr, rp2) and np.array_equal(cc, cc2) and bits_equal(ca, ca2)
        # the ELL program never expands symmetry (Q5): K counts the stored triangle only
        err, d2, ec, ea = hostlib.ell_from_file(bits, p)
        assert err == 0 and d2[4] == n * d2[3] and np.count_nonzero(ea) <= len(a)


def test_fields_gzip_and_comments(tmp_path):
    ri, ci, a = [1, 2, 2], [2, 1, 3], [3.0, -4.0, 5.0]
    p = str(tmp_path / "I.mtx")
    hostlib.write_mtx(p, 2, 3, ri, ci, a, field="integer", comments=("%c1", "% c2", "%"))
    err, dims, ec, ea = hostlib.ell_from_file(32, p)
    assert err == 0 and ea.tolist() == [3.0, 0.0, -4.0, 5.0] and ec.tolist() == [1, 0, 0, 2]
    hostlib.write_mtx(p, 2, 3, ri, ci, a, field="pattern")
    err, dims, ec, ea = hostlib.ell_from_file(64, p)
    assert err == 0 and ea.tolist() == [1.0, 0.0, 1.0, 1.0]
    with open(p, "rb") as f, gzip.open(p + ".gz", "wb") as z:
        z.write(f.read())
    err, dims, ec2, ea2 = hostlib.ell_from_file(64, p + ".gz", gzip=1)
    assert err == 0 and np.array_equal(ec, ec2) and bits_equal(ea, ea2)


BAD = {
    "no banner": ("%MatrixMarket matrix coordinate real general\n1 1 1\n1 1 1.0\n", errno.EINVAL, 1),
    "double space": ("%%MatrixMarket matrix  coordinate real general\n1 1 1\n1 1 1.0\n", errno.EINVAL, 1),
    "array matrix": ("%%MatrixMarket matrix array real general\n1 1\n1.0\n", errno.EINVAL, 2),
    "complex": ("%%MatrixMarket matrix coordinate complex general\n1 1 1\n1 1 1.0 0\n", errno.EINVAL, 1),
    "size line": ("%%MatrixMarket matrix coordinate real general\n2 2\n1 1 1.0\n", errno.EINVAL, 2),
    "tab separator": ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1\t1 1.0\n", errno.EINVAL, 3),
    "missing value": ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 1\n", errno.EINVAL, 3),
    "row out of range": ("%%MatrixMarket matrix coordinate real general\n2 2 1\n3 1 1.0\n", errno.EINVAL, 3),
    "col zero": ("%%MatrixMarket matrix coordinate real general\n2 2 1\n1 0 1.0\n", errno.EINVAL, 3),
    "truncated": ("%%MatrixMarket matrix coordinate real general\n2 2 2\n1 1 1.0\n", -1, 4),
    "long line": ("%%MatrixMarket matrix coordinate real general\n%" + "x" * 5000 + "\n1 1 1\n1 1 1.0\n", errno.EOVERFLOW, 2),
}


@pytest.mark.parametrize("case", sorted(BAD))
def test_reader_rejects_what_the_reference_rejects(tmp_path, case):
    text, want, line = BAD[case]
    p = str(tmp_path / "bad.mtx")
    with open(p, "w") as f:
        f.write(text)
    err, dims, _, _ = hostlib.ell_from_file(32, p)
    assert err == want if want != -1 else err in (-1, 2 ** 32 - 1)
    assert dims[6] + 1 == line          # "path:line:" the programs print


def test_index_width_limits(tmp_path):
    p = str(tmp_path / "big.mtx")
    with open(p, "w") as f:
        f.write("%%MatrixMarket matrix coordinate real general\n2 3000000000 1\n2 3000000000 2.5\n")
    err, _, _, _ = hostlib.ell_from_file(32, p)
    assert err == errno.ERANGE          # does not fit a 32-bit idx_t, like parse_int32_t (ellspmv.c:384-396)
    err, dims, ec, ea = hostlib.ell_from_file(64, p)
    assert err == 0 and dims[1] == 3000000000 and ec.tolist() == [0, 2999999999] and ea.tolist() == [0.0, 2.5]


def test_programs_fail_loudly_without_a_gpu(tmp_path):
    import ellspmv_b200 as E
    if E.device_count() > 0:
        pytest.skip("has a GPU")
    hostlib.build_host()
    g = load_golden("test_mtx")
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]))
    for prog in ("ellspmv", "ellspmv64", "csrspmv", "csrspmv64"):
        r = subprocess.run([os.path.join(hostlib.BIN, prog), p], capture_output=True, text=True)
        assert r.returncode == 1 and r.stdout == "" and "No such device" in r.stderr
    r = subprocess.run([os.path.join(hostlib.BIN, "ellspmv"), "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("Usage: ellspmv [OPTION..] A [x] [y]")
    r = subprocess.run([os.path.join(hostlib.BIN, "ellspmv")], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("Usage:")
    r = subprocess.run([os.path.join(hostlib.BIN, "ellspmv"), "--repeat=x", p], capture_output=True, text=True)
    assert r.returncode == 1 and "--repeat=x" in r.stderr


def _big_mtx(path, nr, nc, nnz, seed=1, field="real", tail_newline=True, mutate=None):
    rng = np.random.default_rng(seed)
    ri = rng.integers(1, nr + 1, nnz)
    ci = rng.integers(1, nc + 1, nnz)
    a = rng.standard_normal(nnz)
    lines = [f"{r} {c} {v:.17g}" if field == "real" else f"{r} {c}" for r, c, v in zip(ri, ci, a)]
    if mutate:
        mutate(lines)
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} general\n% big\n{nr} {nc} {nnz}\n")
        f.write("\n".join(lines))
        if tail_newline:
            f.write("\n")
    return ri, ci, a


@pytest.mark.parametrize("field,tail_newline", [("real", True), ("real", False), ("pattern", True)])
def test_parallel_reader_equals_serial_reader(tmp_path, monkeypatch, field, tail_newline):
    """>= 65536 entries take the mmap + OpenMP parser; it must give the serial
    reader's arrays, and the serial reader's are pinned by the goldens above."""
    p = str(tmp_path / "big.mtx")
    ri, ci, a = _big_mtx(p, 5000, 7000, 150000, field=field, tail_newline=tail_newline)
    monkeypatch.delenv("ELLSPMV_SERIAL_READER", raising=False)
    err, dims, ec, ea = hostlib.ell_from_file(64, p)
    monkeypatch.setenv("ELLSPMV_SERIAL_READER", "1")
    err2, dims2, ec2, ea2 = hostlib.ell_from_file(64, p)
    assert err == 0 and err2 == 0 and dims == dims2
    assert np.array_equal(ec, ec2) and bits_equal(ea, ea2)
    # and against an independent conversion of the values that were written
    K = dims[3]
    cnt = np.bincount(ri - 1, minlength=5000)
    assert K == cnt.max() and dims[6] == 3 + 150000
    row0 = np.flatnonzero(ri == ri[0])
    want_vals = a[row0] if field == "real" else np.ones(len(row0))
    r = ri[0] - 1
    assert bits_equal(ea[r * K:r * K + len(row0)], want_vals)
    assert np.array_equal(ec[r * K:r * K + len(row0)], ci[row0] - 1)


@pytest.mark.parametrize("what", ["bad index", "tab", "short file", "leading blank", "plus sign", "huge value"])
def test_parallel_reader_falls_back_to_the_serial_answer(tmp_path, monkeypatch, what):
    p = str(tmp_path / "big.mtx")

    def mutate(lines):
        if what == "bad index":
            lines[90000] = "999999 1 1.0"
        elif what == "tab":
            lines[90000] = "1\t1 1.0"
        elif what == "short file":
            del lines[100000:]
        elif what == "leading blank":
            lines[70000] = " 3 4 2.5"          # strtoll skips blanks: the reference accepts this line
        elif what == "plus sign":
            lines[70000] = "+3 4 +2.5"
        elif what == "huge value":
            lines[70000] = "3 4 1e999"
    _big_mtx(p, 5000, 7000, 120000, mutate=mutate)
    monkeypatch.delenv("ELLSPMV_SERIAL_READER", raising=False)
    par = hostlib.ell_from_file(32, p)
    monkeypatch.setenv("ELLSPMV_SERIAL_READER", "1")
    ser = hostlib.ell_from_file(32, p)
    assert par[0] == ser[0] and par[1] == ser[1]          # same error code, same "line" for the message
    if ser[0] == 0:
        assert np.array_equal(par[2], ser[2]) and bits_equal(par[3], ser[3])
    else:
        assert what in ("bad index", "tab", "short file", "huge value")


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_host_sort_rows_matches_reference(tmp_path, name, bits):
    import ctypes as C
    g = load_golden(name)
    e = g[f"idx{bits}"]
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]))
    lib = hostlib.hostlib(bits)
    it = C.c_int32 if bits == 32 else C.c_int64
    dt = np.int32 if bits == 32 else np.int64
    dims = (C.c_int64 * 7)()
    rowptr, colidx, a, ad = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.host_csr_from_file_sd(p.encode(), 0, 2, dims, C.byref(rowptr), C.byref(colidx), C.byref(a), C.byref(ad)) == 0
    assert hostlib._take(lib, rowptr, g["num_rows"] + 1, C.c_int64, np.int64).tolist() == e["rowptr"]
    assert hostlib._take(lib, colidx, dims[3], it, dt).tolist() == e["csrcolidx_sorted"]
    assert bits_equal(hostlib._take(lib, a, dims[3], C.c_double, np.float64), unhex(e["csra_sorted"]))


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_host_ell_sort_rows_equals_sorted_csr_rows(tmp_path, name, bits):
    """Intended ELL --sort-rows: row i == the reference-sorted CSR row i, then padding."""
    import ctypes as C
    g = load_golden(name)
    e = g[f"idx{bits}"]
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]))
    lib = hostlib.hostlib(bits)
    it = C.c_int32 if bits == 32 else C.c_int64
    dt = np.int32 if bits == 32 else np.int64
    dims = (C.c_int64 * 7)()
    colidx, a, ad = C.c_void_p(), C.c_void_p(), C.c_void_p()
    assert lib.host_ell_from_file_sd(p.encode(), 0, 2, dims, C.byref(colidx), C.byref(a), C.byref(ad)) == 0
    K, nr = dims[3], g["num_rows"]
    ec = hostlib._take(lib, colidx, dims[4], it, dt).reshape(nr, K) if K else np.zeros((nr, 0), dtype=dt)
    ea = hostlib._take(lib, a, dims[4], C.c_double, np.float64).reshape(nr, K) if K else np.zeros((nr, 0))
    rowptr = e["rowptr"]
    sc, sa = np.array(e["csrcolidx_sorted"], dtype=dt), unhex(e["csra_sorted"])
    for i in range(nr):
        n = rowptr[i + 1] - rowptr[i]
        assert np.array_equal(ec[i, :n], sc[rowptr[i]:rowptr[i + 1]]) and bits_equal(ea[i, :n], sa[rowptr[i]:rowptr[i + 1]])
        assert np.all(ec[i, n:] == min(i, g["num_columns"] - 1)) and np.all(ea[i, n:] == 0.0)
