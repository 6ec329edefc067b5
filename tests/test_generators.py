"""Synthetic matrices of BASELINE.json's shapes: the direct ELL generator
agrees with pushing the canonical COO stream through ell_from_coo, shard
ranges tile the full matrix, and known row sums hold."""
import numpy as np
import pytest

from conftest import bits_equal

CASES = [("laplace2d", (7, 11), (4.0, -1.0)), ("laplace2d", (1, 6), (4.0, -1.0)), ("laplace2d", (6, 1), (4.0, -1.0)),
         ("stencil27", (4, 5, 3), (26.0, -1.0)), ("stencil27", (1, 1, 1), (0.5, -1.0 / 52)), ("stencil27", (2, 2, 2), (26.0, -1.0)),
         ("random", (37, 23, 6), (0, 0)), ("random", (10, 1000, 32), (0, 0))]


@pytest.mark.parametrize("kind,dims,vals", CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_direct_equals_coo_route(oracle, kind, dims, vals, bits):
    rows, ncols, ri, ci, a = oracle.gen_coo(kind, dims, vals, seed=42, bits=bits)
    K, ellsize, _, ec, ea = oracle.ell_from_coo(rows, ncols, ri, ci, a)
    K2, ncols2, ec2, ea2, real = oracle.gen_ell(kind, dims, vals, seed=42, bits=bits)
    assert real == len(a) and ncols2 == ncols
    if kind == "random" or min(dims[: 2 if kind == "laplace2d" else 3]) >= 3:
        assert K == K2
    if K == K2:
        assert np.array_equal(ec, ec2) and bits_equal(ea, ea2)
    else:
        # tiny grids: no row reaches the full stencil, ell_from_coo picks a smaller K;
        # the entries in front of the padding must still agree
        e1, e2 = ec.reshape(rows, K), ec2.reshape(rows, K2)
        assert np.array_equal(e1, e2[:, :K]) and bits_equal(ea.reshape(rows, K), ea2.reshape(rows, K2)[:, :K])
        assert np.all(ea2.reshape(rows, K2)[:, K:] == 0.0)


@pytest.mark.parametrize("kind,dims,vals", CASES)
def test_shards_tile_the_matrix(oracle, kind, dims, vals):
    K, ncols, ec, ea, _ = oracle.gen_ell(kind, dims, vals)
    rows = len(ea) // K
    cuts = sorted({0, rows // 3, (2 * rows) // 3, rows})
    parts_c, parts_a = [], []
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        _, _, c, a, _ = oracle.gen_ell(kind, dims, vals, row_begin=lo, row_end=hi)
        parts_c.append(c)
        parts_a.append(a)
    assert np.array_equal(np.concatenate(parts_c), ec) and bits_equal(np.concatenate(parts_a), ea)


def test_laplacian_times_ones_is_the_boundary_indicator(oracle):
    nx, ny = 9, 7
    K, ncols, ec, ea, _ = oracle.gen_ell("laplace2d", (nx, ny), (4.0, -1.0))
    y = np.zeros(nx * ny)
    oracle.ellgemv(nx * ny, y, np.ones(ncols), K, ec, ea)
    i, j = np.divmod(np.arange(nx * ny), ny)
    missing = (i == 0).astype(float) + (i == nx - 1) + (j == 0) + (j == ny - 1)
    assert np.array_equal(y, missing)


def test_random_generator_definition(oracle):
    """col = mulhi64(splitmix64(seed ^ (r*K+l)), ncols); val in [-1, 1)."""
    rows, ncols, K, seed = 5, 1000, 3, 42
    _, _, ec, ea, _ = oracle.gen_ell("random", (rows, ncols, K), seed=seed)
    sm = oracle.lib.oracle_splitmix64_export
    for e in range(rows * K):
        u = sm(seed ^ e)
        assert ec[e] == (u * ncols) >> 64
        assert ea[e] == 2.0 * ((sm(u) >> 11) * 2.0 ** -53) - 1.0
    assert np.all((ea >= -1.0) & (ea < 1.0))
