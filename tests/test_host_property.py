"""Property test (hypothesis): for arbitrary small COO input the host programs'
reader + converters give exactly the oracle's ELL and CSR arrays (which the
goldens tie to the unmodified reference), in both index widths, with and
without the diagonal split, and --sort-rows orders every row."""
import os
import tempfile

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

import hostlib
from conftest import bits_equal
from oracle.pyoracle import Oracle

ORC = Oracle()


@st.composite
def coo(draw):
    nr = draw(st.integers(1, 12))
    nc = draw(st.integers(1, 12))
    nnz = draw(st.integers(0, 60))
    ri = draw(st.lists(st.integers(1, nr), min_size=nnz, max_size=nnz))
    ci = draw(st.lists(st.integers(1, nc), min_size=nnz, max_size=nnz))
    vals = draw(st.lists(st.floats(-1e3, 1e3, allow_nan=False, width=64) | st.sampled_from([0.0, -0.0, 1e-310, 1e300]),
                         min_size=nnz, max_size=nnz))
    return nr, nc, ri, ci, vals


@settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(coo(), st.sampled_from([32, 64]))
def test_host_conversion_equals_oracle(case, bits):
    import ctypes as C
    nr, nc, ri, ci, vals = case
    dt = np.int32 if bits == 32 else np.int64
    r, c, a = np.array(ri, dtype=dt), np.array(ci, dtype=dt), np.array(vals, dtype=np.float64)
    with tempfile.TemporaryDirectory() as tmp:
        p = os.path.join(tmp, "A.mtx")
        hostlib.write_mtx(p, nr, nc, ri, ci, a, comments=())
        err, dims, ec, ea = hostlib.ell_from_file(bits, p)
        K, ellsize, diagsize, ec2, ea2 = ORC.ell_from_coo(nr, nc, r, c, a)
        assert err == 0 and dims[3:6] == [K, ellsize, diagsize]
        assert np.array_equal(ec, ec2) and bits_equal(ea, ea2)
        err, dims, rp, cc, ca = hostlib.csr_from_file(bits, p)
        rp2, cc2, ca2, lo, hi = ORC.csr_from_coo(nr, nc, r, c, a)
        assert err == 0 and dims[3:6] == [len(a), lo, hi]
        assert np.array_equal(rp, rp2) and np.array_equal(cc, cc2) and bits_equal(ca, ca2)
        # --sort-rows: same multiset per row, ascending columns, and exactly the oracle's tie order
        lib = hostlib.hostlib(bits)
        it = C.c_int32 if bits == 32 else C.c_int64
        d = (C.c_int64 * 7)()
        prp, pc, pa, pad = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        assert lib.host_csr_from_file_sd(p.encode(), 0, 2, d, C.byref(prp), C.byref(pc), C.byref(pa), C.byref(pad)) == 0
        hostlib._take(lib, prp, nr + 1, C.c_int64, np.int64)
        sc = hostlib._take(lib, pc, d[3], it, dt)
        sa = hostlib._take(lib, pa, d[3], C.c_double, np.float64)
        ORC.rowsort(nr, rp2, cc2, ca2)
        assert np.array_equal(sc, cc2) and bits_equal(sa, ca2)
        for i in range(nr):
            assert np.all(np.diff(sc[rp2[i]:rp2[i + 1]]) >= 0)
        if nr <= nc:
            # diagonal split (declared-order semantics for ELL)
            colidx, av, ad = C.c_void_p(), C.c_void_p(), C.c_void_p()
            assert lib.host_ell_from_file_sd(p.encode(), 0, 1, d, C.byref(colidx), C.byref(av), C.byref(ad)) == 0
            K3, ellsize3, diag3, ec3, ea3, ad3 = ORC.ell_from_coo_sd(nr, nc, r, c, a)
            assert list(d)[3:6] == [K3, ellsize3, diag3]
            assert np.array_equal(hostlib._take(lib, colidx, d[4], it, dt), ec3)
            assert bits_equal(hostlib._take(lib, av, d[4], C.c_double, np.float64), ea3)
            assert bits_equal(hostlib._take(lib, ad, d[5], C.c_double, np.float64), ad3)
