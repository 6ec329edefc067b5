"""GPU parity of the CSR comparison path (csrspmv_cuda_*) against the
oracle's csrgemv and the golden vectors of the unmodified reference."""
import numpy as np
import pytest

import ellspmv_b200 as E
from conftest import GOLDEN_CASES, bits_equal, load_golden, unhex

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_golden_reference_vectors(lib, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    dt = np.int32 if bits == 32 else np.int64
    rowptr = np.array(e["rowptr"], dtype=np.int64)
    cc, ca = np.array(e["csrcolidx"], dtype=dt), unhex(e["csra"])
    A = E.CsrMatrix.upload(g["num_rows"], g["num_columns"], rowptr, cc, ca)
    y = unhex(g["y0"])
    A.spmv(y, unhex(g["x"]), 1, E.ACCUMULATE)
    A.free()
    assert bits_equal(y, unhex(e["y_csr"]))


def ragged_csr(rng, nr, nc, maxlen, dt):
    lens = rng.integers(0, maxlen + 1, nr)
    lens[rng.integers(0, nr, max(nr // 10, 1))] = 0
    if nr > 3:
        lens[nr // 2] = 5 * maxlen + 3000          # one very long row (spans several tiles)
    rowptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    return rowptr, rng.integers(0, nc, nnz).astype(dt), rng.standard_normal(nnz)


@pytest.mark.parametrize("shape", [(1, 1, 1), (100, 50, 4), (129, 1000, 40), (5000, 5000, 9), (1025, 300, 70)])
@pytest.mark.parametrize("bits", [32, 64])
def test_bit_exact_vs_oracle(lib, oracle, shape, bits):
    nr, nc, maxlen = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + nc + maxlen + bits)
    rowptr, cc, ca = ragged_csr(rng, nr, nc, maxlen, dt)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want = y0.copy()
    oracle.csrgemv(nr, want, x, rowptr, cc, ca)
    want0 = np.zeros(nr)
    oracle.csrgemv(nr, want0, x, rowptr, cc, ca)
    for kflag in (0, E.KERNEL_THREAD, 3):     # auto / smem-staged stream / thread-per-row scalar: all bit-exact
        A = E.CsrMatrix.upload(nr, nc, rowptr, cc, ca, kflag)
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        assert bits_equal(y, want), kflag
        A.spmv(y, x, 2, E.OVERWRITE)
        assert bits_equal(y, want0), kflag
        A.free()
    # tolerance modes
    absprod = np.zeros(nr)
    np.add.at(absprod, np.repeat(np.arange(nr), np.diff(rowptr)), np.abs(ca * x[cc]))
    bound = (np.diff(rowptr) + 2) * 2.0 ** -53 * absprod
    for flags in (E.FMA, E.KERNEL_WARP, E.KERNEL_WARP | E.FMA):
        A = E.CsrMatrix.upload(nr, nc, rowptr, cc, ca, flags)
        y = np.zeros(nr)
        A.spmv(y, x, 1, E.ACCUMULATE)
        A.free()
        assert np.all(np.abs(y - want0) <= bound + 1e-300), flags


def test_reference_shaped_operator(lib):
    g = load_golden("test_mtx")
    e = g["idx32"]
    rowptr = np.array(e["rowptr"], dtype=np.int64)
    cc, ca = np.array(e["csrcolidx"], dtype=np.int32), unhex(e["csra"])
    y, x = np.zeros(4), np.ones(5)
    assert E.csrgemv(4, y, 5, x, 9, 1, 5, rowptr, cc, ca) == 0
    assert y.tolist() == [3, 1, 3, 6]


def test_random_csr_is_the_ell_matrix(lib, oracle):
    """BASELINE config 4: the CSR view of the random matrix has the same
    entries as the ELL one, so both paths must give identical bits."""
    dims = (30000, 30000, 32)
    x = np.random.default_rng(2).standard_normal(dims[1])
    Ae = E.EllMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32)
    Ac = E.CsrMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32)
    ye, yc = np.zeros(dims[0]), np.zeros(dims[0])
    Ae.spmv(ye, x, 1, E.ACCUMULATE)
    Ac.spmv(yc, x, 1, E.ACCUMULATE)
    Ae.free(); Ac.free()
    K, ncols, ec, ea, _ = oracle.gen_ell("random", dims, seed=42)
    want = np.zeros(dims[0])
    oracle.ellgemv(dims[0], want, x, K, ec, ea)
    assert bits_equal(ye, want) and bits_equal(yc, want)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_sort_rows_program_and_kernel(tmp_path, lib, name):
    """csrspmv --sort-rows: stdout equals the reference program's, and the
    kernel on the reference-sorted arrays gives the reference's y."""
    import os
    import subprocess

    import hostlib
    g = load_golden(name)
    e = g["idx32"]
    A = E.CsrMatrix.upload(g["num_rows"], g["num_columns"], np.array(e["rowptr"], dtype=np.int64),
                           np.array(e["csrcolidx_sorted"], dtype=np.int32), unhex(e["csra_sorted"]))
    y = unhex(g["y0"])
    A.spmv(y, unhex(g["x"]), 1, E.ACCUMULATE)
    A.free()
    assert bits_equal(y, unhex(e["y_csr_sorted"]))
    hostlib.build_host()
    p = str(tmp_path / "A.mtx")
    hostlib.write_mtx(p, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]), comments=())
    r = subprocess.run([os.path.join(hostlib.BIN, "csrspmv"), "--sort-rows", p], capture_output=True, text=True,
                       env=dict(os.environ, LC_ALL="C"))
    assert r.returncode == 0 and r.stdout == g["program"]["csrspmv_sorted"]["stdout"]


def test_upload_validates_rowptr_and_columns(lib):
    """csrspmv_cuda_upload checks what every ELL upload checks: a decreasing rowptr or a column
    outside [0, num_columns) is EINVAL, not an out-of-bounds gather."""
    rowptr = np.array([0, 2, 4, 6], dtype=np.int64)
    a = np.ones(6)
    for dt in (np.int32, np.int64):
        ok = np.array([0, 1, 2, 3, 4, 5], dtype=dt)
        E.CsrMatrix.upload(3, 6, rowptr, ok, a).free()
        for bad in ([0, 1, 2, 6, 4, 5], [0, 1, -1, 3, 4, 5], [-2] * 6):
            with pytest.raises(E.EllspmvCudaError) as ei:
                E.CsrMatrix.upload(3, 6, rowptr, np.array(bad, dtype=dt), a)
            assert ei.value.errno == 22
        with pytest.raises(E.EllspmvCudaError) as ei:
            E.CsrMatrix.upload(3, 6, np.array([0, 4, 2, 6], dtype=np.int64), ok, a)
        assert ei.value.errno == 22


def test_host_call_reads_only_the_referenced_x_range(lib, oracle):
    rng = np.random.default_rng(8)
    nr, nc = 300, 5000
    rowptr, ec, ea = ragged_csr(rng, nr, 400, 9, np.int32)
    ec = (ec + 1000).astype(np.int32)                  # columns 1000..1399 only
    x = rng.standard_normal(nc)
    want = np.zeros(nr)
    oracle.csrgemv(nr, want, x, rowptr, ec, ea)
    xp = np.full(nc, np.nan)
    xp[ec.min():ec.max() + 1] = x[ec.min():ec.max() + 1]
    A = E.CsrMatrix.upload(nr, nc, rowptr, ec, ea)
    y = np.zeros(nr)
    A.spmv(y, xp, 1, E.ACCUMULATE)
    A.free()
    assert bits_equal(y, want)


def balanced_csr(rng, nr, nc, K, dt, uniform):
    lens = np.full(nr, K) if uniform else rng.integers(max(K - K // 5, 0), K + 1, nr)
    if not uniform:
        lens[rng.integers(0, nr)] = K             # the longest row sets the width
    rowptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    return rowptr, rng.integers(0, nc, nnz).astype(dt), rng.standard_normal(nnz)


@pytest.mark.parametrize("shape", [(1, 7, 3), (300, 200, 5), (5000, 5000, 9), (4097, 900, 32), (1500, 100000, 40)])
@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("uniform", [False, True])
def test_auto_runs_balanced_rows_through_the_ell_view(lib, oracle, shape, bits, uniform):
    """KERNEL_AUTO keeps a sliced-ELL view (per-row lengths) of a CSR matrix whose rows are balanced and
    launches the ELL kernels: same bits as csrgemv, also when x is non-finite exactly where the view
    has padded slots (they never enter the arithmetic)."""
    nr, nc, K = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr * 31 + K + bits + uniform)
    rowptr, ec, ea = balanced_csr(rng, nr, nc, K, dt, uniform)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    A = E.CsrMatrix.upload(nr, nc, rowptr, ec, ea)
    i = A.info()
    assert i.ell_view == (2 if uniform or nr == 1 else 1) and i.max_row_len == K, (i.ell_view, i.max_row_len)
    assert "ELL view" in A.describe()
    for xv in (x, np.where(rng.random(nc) < 0.02, np.inf, x), np.where(rng.random(nc) < 0.02, np.nan, x)):
        want = y0.copy()
        for _ in range(2):
            oracle.csrgemv(nr, want, xv, rowptr, ec, ea)
        y = y0.copy()
        A.spmv(y, xv, 2, E.ACCUMULATE)
        if np.isnan(xv).any():
            # NaN payloads are not propagated identically: same rows NaN, the rest bit-equal
            assert np.array_equal(np.isnan(y), np.isnan(want))
            ok = ~np.isnan(want)
            assert bits_equal(y[ok], want[ok])
        else:
            assert bits_equal(y, want)
    r2, c2, a2 = A.download(nr, int(rowptr[-1]), bits)       # the CSR arrays themselves stay as uploaded
    assert np.array_equal(r2, rowptr) and np.array_equal(c2, ec) and bits_equal(a2, ea)
    A.free()
    # the native kernels on the same matrix give the same bits
    for kern in (E.KERNEL_THREAD, E.KERNEL_CSR_SCALAR):
        B = E.CsrMatrix.upload(nr, nc, rowptr, ec, ea, kern)
        assert B.info().ell_view == 0
        y = y0.copy()
        B.spmv(y, x, 1, E.ACCUMULATE)
        w = y0.copy()
        oracle.csrgemv(nr, w, x, rowptr, ec, ea)
        assert bits_equal(y, w)
        B.free()


@pytest.mark.parametrize("mode", [E.ACCUMULATE, E.OVERWRITE])
def test_pipelined_host_call_through_the_ell_view(lib, oracle, mode):
    """One launch on >= 2^20 balanced CSR rows with host vectors takes the ELL host call's pipeline
    (row chunks uploaded, run and downloaded on three streams) on the view: same bits as csrgemv,
    pageable and pinned vectors, with and without the separately stored diagonal, and the plain
    path (repeat = 2) agrees."""
    import torch
    rng = np.random.default_rng(17)
    nr = nc = (1 << 20) + 12345
    rowptr, ec, ea = balanced_csr(rng, nr, nc, 6, np.int32, uniform=False)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want = y0.copy() if mode == E.ACCUMULATE else np.zeros(nr)
    oracle.csrgemv(nr, want, x, rowptr, ec, ea)
    A = E.CsrMatrix.upload(nr, nc, rowptr, ec, ea)
    assert A.info().ell_view == 1
    y = y0.copy()
    secs = A.spmv(y, x, 1, mode)
    assert bits_equal(y, want) and secs[0] > 0
    xp = torch.from_numpy(x).pin_memory()
    yp = torch.from_numpy(y0.copy()).pin_memory()
    A.spmv(yp.numpy(), xp.numpy(), 1, mode)
    assert bits_equal(yp.numpy(), want)
    if mode == E.ACCUMULATE:
        want2 = want.copy()
        oracle.csrgemv(nr, want2, x, rowptr, ec, ea)
        y = y0.copy()
        A.spmv(y, x, 2, mode)                      # repeat = 2: the plain path, vectors resident
        assert bits_equal(y, want2)
    ad = rng.standard_normal(nr)
    w = y0.copy() if mode == E.ACCUMULATE else np.zeros(nr)
    oracle.csrgemvsd(nr, w, x, rowptr, ec, ea, ad)
    A.set_diagonal(ad)
    y = y0.copy()
    A.spmv(y, x, 1, mode)
    assert bits_equal(y, w)
    A.free()


def test_ell_view_padding_never_meets_x(lib, oracle):
    """Every row but one is short; x is +inf on every column that only the view's padded slots touch."""
    rng = np.random.default_rng(5)
    nr, nc, K = 700, 3000, 10
    lens = np.full(nr, 9)
    lens[3] = K
    rowptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    ec = rng.integers(0, nc, nnz).astype(np.int32)
    ea = rng.standard_normal(nnz)
    x = rng.standard_normal(nc)
    last_cols = ec[rowptr[1:] - 1]                 # what the padded slots point at
    ea[rowptr[1:] - 1] = 0.0                       # ... with a real coefficient of exactly 0 there
    x[last_cols] = 1.0                             # finite: 0 * 1 = 0 in csrgemv
    want = np.zeros(nr)
    oracle.csrgemv(nr, want, x, rowptr, ec, ea)
    A = E.CsrMatrix.upload(nr, nc, rowptr, ec, ea)
    assert A.info().ell_view == 1
    y = np.zeros(nr)
    A.spmv(y, x, 1, E.ACCUMULATE)
    assert bits_equal(y, want)
    # now non-finite exactly there: csrgemv's real entry gives 0*inf = NaN once; a view that also
    # multiplied its padded slot would agree by accident, so use -inf/+inf signs to tell: the
    # reference result is NaN either way, equality of NaN-ness and of all other rows is the check
    x[last_cols[::2]] = np.inf
    want = np.zeros(nr)
    oracle.csrgemv(nr, want, x, rowptr, ec, ea)
    y = np.zeros(nr)
    A.spmv(y, x, 1, E.ACCUMULATE)
    assert np.array_equal(np.isnan(y), np.isnan(want))
    ok = ~np.isnan(want)
    assert bits_equal(y[ok], want[ok])
    A.free()


def test_random_generated_csr_uses_the_uniform_view(lib, oracle):
    dims = (6000, 6000, 32)
    K, ncols, ec, ea, _ = oracle.gen_ell("random", dims, seed=42, bits=32)
    x = np.random.default_rng(2).standard_normal(ncols)
    want = np.zeros(dims[0])
    oracle.ellgemv(dims[0], want, x, K, ec, ea)
    A = E.CsrMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32)
    assert A.info().ell_view == 2
    y = np.zeros(dims[0])
    A.spmv(y, x, 1, E.ACCUMULATE)
    A.free()
    assert bits_equal(y, want)


STENCIL_GRIDS = [("laplace2d", (1, 1)), ("laplace2d", (1, 7)), ("laplace2d", (96, 50)), ("laplace2d", (300, 257)),
                 ("stencil27", (1, 1, 1)), ("stencil27", (2, 1, 3)), ("stencil27", (24, 20, 16)), ("stencil27", (40, 36, 33))]


@pytest.mark.parametrize("kind,dims", STENCIL_GRIDS)
@pytest.mark.parametrize("bits", [32, 64])
def test_generated_stencils_in_csr_form(lib, oracle, kind, dims, bits):
    """csrspmv_cuda_generate for the two stencils: the arrays are what csr_from_coo (csrspmv.c:1390-1475)
    makes of the canonical COO stream (no padding, shorter rows at the grid boundary); the SpMV runs
    through the sliced-ELL view with per-row lengths AND offset patterns, and gives csrgemv's bits --
    also with variable coefficients uploaded on the same structure, and with non-finite x."""
    rng = np.random.default_rng(len(dims) * 1000 + int(np.prod(dims)) + bits)
    rows, ncols, ri, ci, a = oracle.gen_coo(kind, dims, (26.0, -1.0), bits=bits)
    rowptr, cc, ca, lo, hi = oracle.csr_from_coo(rows, ncols, ri, ci, a)
    gk = E.GEN_LAPLACE2D if kind == "laplace2d" else E.GEN_STENCIL27
    A = E.CsrMatrix.generate(gk, dims, idx_bits=bits, vals=(26.0, -1.0))
    i = A.info()
    assert (i.num_rows, i.num_columns, i.csrsize) == (rows, ncols, len(ca))
    assert (i.min_row_len, i.max_row_len) == (lo, hi)
    r2, c2, a2 = A.download(rows, len(ca), bits)
    assert np.array_equal(r2, rowptr) and np.array_equal(c2, cc) and bits_equal(a2, ca)
    assert i.ell_view == (2 if lo == hi else 1)
    if rows >= 4000:
        # interior rows (and, with one id per thread, the boundary rows too) are on the dictionary
        assert i.ell_pattern_rows >= 0.6 * rows, (i.ell_pattern_rows, rows)
    x = rng.standard_normal(ncols)
    y0 = rng.standard_normal(rows)
    for xv in (x, np.where(rng.random(ncols) < 0.03, np.inf, x)):
        want = y0.copy()
        oracle.csrgemv(rows, want, xv, rowptr, cc, ca)
        y = y0.copy()
        A.spmv(y, xv, 1, E.ACCUMULATE)
        assert np.array_equal(np.isnan(y), np.isnan(want))
        ok = ~np.isnan(want)
        assert bits_equal(y[ok], want[ok])
    A.free()
    # the same structure with variable coefficients, uploaded: patterns with and without, lanes or group ids
    va = rng.standard_normal(len(ca))
    want = np.zeros(rows)
    oracle.csrgemv(rows, want, x, rowptr, cc, va)
    seen = {}
    for flags in (0, E.NO_PATTERN_LANES, E.NO_PATTERN, E.FMA):
        B = E.CsrMatrix.upload(rows, ncols, rowptr, cc, va, flags)
        seen[flags] = B.info().ell_pattern_rows
        y = np.zeros(rows)
        B.spmv(y, x, 1, E.OVERWRITE)
        if flags & E.FMA:
            assert np.allclose(y, want, rtol=1e-13, atol=1e-13)
        else:
            assert bits_equal(y, want), flags
        B.free()
    assert seen[E.NO_PATTERN] == 0 and seen[0] >= seen[E.NO_PATTERN_LANES]


def powerlaw_csr(rng, nr, nc, dt, min_len=2, max_len=9000, alpha=1.3, empty_frac=0.05):
    """Row lengths ~ Pareto: most rows short, a tail of very long ones."""
    u = rng.random(nr)
    lens = np.minimum((min_len / u ** (1.0 / alpha)).astype(np.int64), max_len)
    lens[rng.random(nr) < empty_frac] = 0
    if nr > 1000:
        lens[nr // 3] = 5000                           # a row of several tiles for the CTA-per-row kernel
        lens[nr // 2] = 257                            # and one just over the threshold
    rowptr = np.zeros(nr + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    return rowptr, rng.integers(0, nc, nnz).astype(dt), rng.standard_normal(nnz)


@pytest.mark.parametrize("shape", [(1, 50), (130, 5000), (4096, 3000), (4097, 70000), (30000, 30000)])
@pytest.mark.parametrize("bits", [32, 64])
def test_auto_runs_skewed_rows_through_sell(lib, oracle, shape, bits):
    """Unbalanced rows (power law): KERNEL_AUTO builds SELL-128-sigma (rows sorted by length in
    windows of 4096, width per slice, rows beyond 4096 entries one CTA each) -- csrgemv's bits,
    with accumulate/overwrite, the separate diagonal, non-finite x, and several launches."""
    nr, nc = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + bits)
    rowptr, ec, ea = powerlaw_csr(rng, nr, nc, dt)
    lens = np.diff(rowptr)
    if nr > 1000:
        assert lens.max() > 1000                       # the CTA-per-row kernel has work, over several tiles
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    A = E.CsrMatrix.upload(nr, nc, rowptr, ec, ea)
    i = A.info()
    if nr >= 130:
        assert i.kernel == E.KERNEL_CSR_SELL and i.ell_view == 0, (i.kernel, i.ell_view)
        T = i.sell_long_len
        assert T == 256
        assert i.sell_real + int(lens[lens > T].sum()) == rowptr[-1]
        assert i.sell_long_rows == int((lens > T).sum())
        if nr >= 4096:
            assert i.sell_slots < 2.2 * i.sell_real + 4096 * 8    # sorting by length keeps the padding bounded
    for xv in (x, np.where(rng.random(nc) < 0.01, np.inf, x)):
        want = y0.copy()
        for _ in range(2):
            oracle.csrgemv(nr, want, xv, rowptr, ec, ea)
        y = y0.copy()
        A.spmv(y, xv, 2, E.ACCUMULATE)
        assert np.array_equal(np.isnan(y), np.isnan(want))
        ok = ~np.isnan(want)
        assert bits_equal(y[ok], want[ok])
    w0 = np.zeros(nr)
    oracle.csrgemv(nr, w0, x, rowptr, ec, ea)
    y = np.full(nr, 3.0)
    A.spmv(y, x, 1, E.OVERWRITE)
    assert bits_equal(y, w0)
    if nr <= nc:
        ad = rng.standard_normal(nr)
        want = y0.copy()
        oracle.csrgemvsd(nr, want, x, rowptr, ec, ea, ad)
        A.set_diagonal(ad)
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        assert bits_equal(y, want)
    A.free()
    # explicit selector, and the native kernels on the same matrix
    for kern in (E.KERNEL_CSR_SELL, E.KERNEL_THREAD, E.KERNEL_CSR_SCALAR):
        B = E.CsrMatrix.upload(nr, nc, rowptr, ec, ea, kern)
        assert B.info().kernel == kern or (kern == E.KERNEL_CSR_SELL and rowptr[-1] == 0)
        y = y0.copy()
        B.spmv(y, x, 1, E.ACCUMULATE)
        w = y0.copy()
        oracle.csrgemv(nr, w, x, rowptr, ec, ea)
        assert bits_equal(y, w), kern
        B.free()
