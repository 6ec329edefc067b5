"""Function-level check of the oracle restatement against the UNMODIFIED
reference compiled into oracle/_ref/ (present in the build container and,
as prebuilt files, on the GPU box; skipped when absent).  Never reads
/root/reference at run time."""
import numpy as np
import pytest

from conftest import bits_equal
from oracle.pyoracle import Reference

pytestmark = pytest.mark.skipif(not Reference.available("ell", 32), reason="oracle/_ref not built")


def random_coo(rng, nr, nc, nnz, dt):
    ri = rng.integers(1, nr + 1, nnz).astype(dt)
    ci = rng.integers(1, nc + 1, nnz).astype(dt)
    a = rng.standard_normal(nnz)
    return ri, ci, a


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("shape", [(200, 200, 1800), (300, 120, 2000), (120, 300, 2000), (64, 64, 64 * 64), (17, 5, 3)])
def test_ell_conversion_and_gemv(oracle, bits, shape):
    nr, nc, nnz = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr * 1000 + nc + bits)
    ri, ci, a = random_coo(rng, nr, nc, nnz, dt)
    ref = Reference("ell", bits)
    K, ellsize, diagsize, ec, ea = ref.ell_from_coo(nr, nc, ri, ci, a)
    K2, ellsize2, diagsize2, ec2, ea2 = oracle.ell_from_coo(nr, nc, ri, ci, a)
    assert (K, ellsize, diagsize) == (K2, ellsize2, diagsize2)
    assert np.array_equal(ec, ec2) and bits_equal(ea, ea2)
    x = rng.standard_normal(nc)
    y_ref = rng.standard_normal(nr)
    y_orc = y_ref.copy()
    ref.ellgemv(nr, y_ref, nc, x, K, ec, ea, repeat=2)
    oracle.ellgemv(nr, y_orc, x, K, ec2, ea2)
    oracle.ellgemv(nr, y_orc, x, K, ec2, ea2)
    assert bits_equal(y_ref, y_orc)


@pytest.mark.parametrize("bits", [32, 64])
@pytest.mark.parametrize("shape", [(200, 200, 1800), (300, 120, 2000), (120, 300, 2000), (17, 5, 3)])
def test_csr_conversion_and_gemv(oracle, bits, shape):
    nr, nc, nnz = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr * 77 + nc + bits)
    ri, ci, a = random_coo(rng, nr, nc, nnz, dt)
    ref = Reference("csr", bits)
    rowptr, cc, ca, lo, hi = ref.csr_from_coo(nr, nc, ri, ci, a)
    rowptr2, cc2, ca2, lo2, hi2 = oracle.csr_from_coo(nr, nc, ri, ci, a)
    assert np.array_equal(rowptr, rowptr2) and np.array_equal(cc, cc2) and bits_equal(ca, ca2)
    assert (lo, hi) == (lo2, hi2)
    x = rng.standard_normal(nc)
    y_ref = rng.standard_normal(nr)
    y_orc = y_ref.copy()
    ref.csrgemv(nr, y_ref, nc, x, rowptr, cc, ca, 1, lo, hi)
    oracle.csrgemv(nr, y_orc, x, rowptr2, cc2, ca2)
    assert bits_equal(y_ref, y_orc)


@pytest.mark.parametrize("kind,dims,vals", [("laplace2d", (13, 9), (4.0, -1.0)),
                                            ("stencil27", (5, 4, 6), (26.0, -1.0)),
                                            ("random", (50, 70, 8), (0.0, 0.0))])
def test_reference_ell_of_synthetic_coo_equals_direct_generator(oracle, kind, dims, vals):
    """ell_from_coo(reference) of the canonical COO stream == the direct ELL generator."""
    for bits in (32, 64):
        rows, ncols, ri, ci, a = oracle.gen_coo(kind, dims, vals, seed=42, bits=bits)
        K, _, _, ec, ea = Reference("ell", bits).ell_from_coo(rows, ncols, ri, ci, a)
        K2, ncols2, ec2, ea2, real = oracle.gen_ell(kind, dims, vals, seed=42, bits=bits)
        assert (K, ncols) == (K2, ncols2) and real == len(a)
        assert np.array_equal(ec, ec2) and bits_equal(ea, ea2)
