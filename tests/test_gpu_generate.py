"""On-device generators of BASELINE.json's synthetic shapes are bit-identical
to the oracle's (which tests/test_generators.py ties to ell_from_coo), and
full-size properties of the headline workload hold."""
import numpy as np
import pytest

import ellspmv_b200 as E
from conftest import bits_equal

pytestmark = pytest.mark.gpu

KINDS = {"laplace2d": E.GEN_LAPLACE2D, "stencil27": E.GEN_STENCIL27, "random": E.GEN_RANDOM}
CASES = [("laplace2d", (37, 29), (4.0, -1.0)), ("laplace2d", (1, 5), (4.0, -1.0)),
         ("stencil27", (9, 7, 8), (26.0, -1.0)), ("stencil27", (2, 1, 3), (0.5, -1.0 / 52)),
         ("random", (1234, 999, 32), (0.0, 0.0)), ("random", (100, 3_000_000_000, 5), (0.0, 0.0))]


@pytest.mark.parametrize("kind,dims,vals", CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_device_generator_equals_oracle(lib, oracle, kind, dims, vals, bits):
    if bits == 32 and max(dims) > 2 ** 31 - 1:
        with pytest.raises(E.EllspmvCudaError):
            E.EllMatrix.generate(KINDS[kind], dims, vals, 42, bits)
        return
    K, ncols, ec, ea, _ = oracle.gen_ell(kind, dims, vals, seed=42, bits=bits)
    rows = len(ea) // K
    for R in (1, 2, 4):
        A = E.EllMatrix.generate(KINDS[kind], dims, vals, 42, bits, flags=E.rows_per_thread(R))
        i = A.info()
        assert (i.num_rows, i.num_columns, i.rowsize) == (rows, ncols, K)
        assert (i.min_col, i.max_col) == (ec.min(), ec.max())
        c2, a2 = A.download()
        assert np.array_equal(c2, ec) and bits_equal(a2, ea)
        A.free()
    # a shard in the middle
    lo, hi = rows // 3, (2 * rows) // 3 + 1
    A = E.EllMatrix.generate(KINDS[kind], dims, vals, 42, bits, row_begin=lo, row_end=hi)
    c2, a2 = A.download()
    assert np.array_equal(c2, ec[lo * K:hi * K]) and bits_equal(a2, ea[lo * K:hi * K])
    A.free()


def test_generated_matrix_spmv_vs_oracle(lib, oracle):
    for kind, dims, vals, bits in [("laplace2d", (300, 211), (4.0, -1.0), 32),
                                   ("stencil27", (30, 31, 29), (26.0, -1.0), 64),
                                   ("random", (20000, 20000, 32), (0, 0), 32)]:
        K, ncols, ec, ea, _ = oracle.gen_ell(kind, dims, vals, bits=bits)
        rows = len(ea) // K
        x = np.random.default_rng(1).standard_normal(ncols)
        want = np.zeros(rows)
        oracle.ellgemv(rows, want, x, K, ec, ea)
        A = E.EllMatrix.generate(KINDS[kind], dims, vals, 42, bits)
        y = np.zeros(rows)
        A.spmv(y, x, 1, E.ACCUMULATE)
        A.free()
        assert bits_equal(y, want), kind


def test_full_size_laplacian_properties(lib):
    """BASELINE config 2 at full size (8192^2 grid, 67M rows): size-independent
    properties instead of an oracle run.  A*ones is the exact boundary
    indicator (integer arithmetic in fp64), repeat accumulates linearly, and
    A*(2x) == 2*(A*x) bit for bit (scaling by 2 is exact)."""
    import torch
    n = 8192
    A = E.EllMatrix.generate(E.GEN_LAPLACE2D, (n, n), (4.0, -1.0), 42, 32)
    rows = n * n
    s = torch.cuda.current_stream().cuda_stream
    x = torch.ones(rows, dtype=torch.float64, device="cuda")
    y = torch.zeros(rows, dtype=torch.float64, device="cuda")
    A.spmv_device(y, x, E.ACCUMULATE, s)
    g = y.view(n, n)
    want = torch.zeros(n, n, dtype=torch.float64, device="cuda")
    want[0, :] += 1; want[-1, :] += 1; want[:, 0] += 1; want[:, -1] += 1
    assert torch.equal(g, want)
    A.spmv_device(y, x, E.ACCUMULATE, s)
    A.spmv_device(y, x, E.ACCUMULATE, s)
    assert torch.equal(g, 3 * want)
    del want, g
    gen = torch.Generator(device="cuda").manual_seed(7)
    xr = torch.randn(rows, dtype=torch.float64, device="cuda", generator=gen)
    y1 = torch.empty_like(xr)
    A.spmv_device(y1, xr, E.OVERWRITE, s)
    xr *= 2
    A.spmv_device(y, xr, E.OVERWRITE, s)
    assert torch.equal(y, 2 * y1)
    # interior row, checked by hand against the 5-point formula in the oracle's order
    xr /= 2
    r = 4000 * n + 4000
    xs = xr[[r - n, r - 1, r, r + 1, r + n]].cpu().numpy()
    acc = 0.0
    for v, xv in zip([-1.0, -1.0, 4.0, -1.0, -1.0], xs):
        acc = acc + v * xv
    assert y1[r].item() == 0.0 + acc
    torch.cuda.synchronize()
    A.free()


def test_full_size_stencil27_properties(lib):
    """BASELINE config 3 at full size (27-point 384^3, IDXTYPEWIDTH=64, 56.6M rows,
    24 GB matrix): A*ones = 26 - (#neighbours) exactly (small integers in fp64),
    the default index-narrowed device layout and the wide one give the same bits, and scaling x by 2
    scales y by 2 bit for bit."""
    import torch
    n = 384
    rows = n ** 3
    s = torch.cuda.current_stream().cuda_stream
    x = torch.ones(rows, dtype=torch.float64, device="cuda")
    c = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    c[0] = c[-1] = 2.0
    want = 27.0 - (c[:, None, None] * c[None, :, None] * c[None, None, :]).reshape(-1)   # 26 - (count - 1)
    ys = []
    for flags in (E.WIDE_INDEX, 0):
        A = E.EllMatrix.generate(E.GEN_STENCIL27, (n, n, n), (26.0, -1.0), 42, 64, flags=flags)
        i = A.info()
        assert (i.num_rows, i.rowsize, i.idx_width_bits, i.dev_idx_bits) == (rows, 27, 64, 64 if flags else 32)
        y = torch.zeros(rows, dtype=torch.float64, device="cuda")
        A.spmv_device(y, x, E.ACCUMULATE, s)
        assert torch.equal(y, want)
        gen = torch.Generator(device="cuda").manual_seed(3)
        xr = torch.randn(rows, dtype=torch.float64, device="cuda", generator=gen)
        A.spmv_device(y, xr, E.OVERWRITE, s)
        y2 = torch.empty_like(y)
        xr *= 2
        A.spmv_device(y2, xr, E.OVERWRITE, s)
        assert torch.equal(y2, 2 * y)
        ys.append(y)
        torch.cuda.synchronize()
        A.free()
        del y2, xr
    assert torch.equal(ys[0], ys[1])


@pytest.mark.parametrize("rank", [0, 3])
def test_full_size_config5_shard_properties(lib, rank):
    """BASELINE config 5 (27-point 768^3, IDXTYPEWIDTH=64) as one of its 8 row shards at full size
    (56.6M rows of 453M, global column indices, x of full length -- what rank `rank` of
    `bench.py --gpus 8` holds): A*ones = 27 - (#stencil points) exactly, the indices are narrowed
    (453M columns < 2^31) and mostly taken from offset patterns, and scaling x by 2 scales y by 2 bit
    for bit.  The exchange itself is covered by test_gpu_sharded / bench.py's parity_check."""
    import torch
    n, world = 768, 8
    rows = n ** 3
    lo, hi = rank * (rows // world), (rank + 1) * (rows // world)
    s = torch.cuda.current_stream().cuda_stream
    A = E.EllMatrix.generate(E.GEN_STENCIL27, (n, n, n), (26.0, -1.0), 42, 64, row_begin=lo, row_end=hi)
    i = A.info()
    assert (i.num_rows, i.num_columns, i.rowsize, i.idx_width_bits, i.dev_idx_bits) == (hi - lo, rows, 27, 64, 32)
    assert i.pattern_rows >= 0.9 * (hi - lo)
    c = torch.full((n,), 3.0, dtype=torch.float64, device="cuda")
    c[0] = c[-1] = 2.0
    planes = n // world
    want = 27.0 - (c[rank * planes:(rank + 1) * planes, None, None] * c[None, :, None] * c[None, None, :]).reshape(-1)
    x = torch.ones(rows, dtype=torch.float64, device="cuda")
    y = torch.zeros(hi - lo, dtype=torch.float64, device="cuda")
    A.spmv_device(y, x, E.ACCUMULATE, s)
    assert torch.equal(y, want)
    del want
    gen = torch.Generator(device="cuda").manual_seed(11 + rank)
    x.normal_(generator=gen)
    A.spmv_device(y, x, E.OVERWRITE, s)
    y2 = torch.empty_like(y)
    x *= 2
    A.spmv_device(y2, x, E.OVERWRITE, s)
    assert torch.equal(y2, 2 * y)
    # the same rows without patterns and with the caller's 64-bit indices: same bits
    torch.cuda.synchronize()
    A.free()
    B = E.EllMatrix.generate(E.GEN_STENCIL27, (n, n, n), (26.0, -1.0), 42, 64, row_begin=lo, row_end=hi,
                             flags=E.WIDE_INDEX | E.NO_PATTERN)
    assert B.info().dev_idx_bits == 64 and B.info().pattern_rows == 0
    y3 = torch.empty_like(y)
    B.spmv_device(y3, x, E.OVERWRITE, s)
    assert torch.equal(y3, y2)
    torch.cuda.synchronize()
    B.free()


def test_full_size_random_ell_equals_csr(lib):
    """BASELINE config 4 at full size (50M x 32, 19 GB per format): the ELL
    path and the CSR comparison path hold the same entries in the same order,
    so y must agree bit for bit; spot rows are re-derived on the host from the
    generator's definition."""
    import torch
    dims = (50_000_000, 50_000_000, 32)
    s = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(dims[1], dtype=torch.float64, device="cuda", generator=gen)
    Ae = E.EllMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32)
    ye = torch.zeros(dims[0], dtype=torch.float64, device="cuda")
    Ae.spmv_device(ye, x, E.ACCUMULATE, s)
    torch.cuda.synchronize()
    Ae.free()
    # the staged-gather path (column blocks, gather staged through HBM) must give the same bits
    As = E.EllMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32, flags=E.STAGED_GATHER)
    ysg = torch.zeros(dims[0], dtype=torch.float64, device="cuda")
    As.spmv_device(ysg, x, E.ACCUMULATE, s)
    torch.cuda.synchronize()
    assert torch.equal(ye, ysg)
    As.spmv_device(ysg, x, E.OVERWRITE, s)
    torch.cuda.synchronize()
    assert torch.equal(ye, ysg)
    As.free()
    del ysg
    Ac = E.CsrMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32)
    yc = torch.zeros(dims[0], dtype=torch.float64, device="cuda")
    Ac.spmv_device(yc, x, E.ACCUMULATE, s)
    torch.cuda.synchronize()
    Ac.free()
    assert torch.equal(ye, yc)
    # spot rows from the definition: u = splitmix64(seed ^ (r*K+l)), col = mulhi(u, ncols), val = 2*(sm(u)>>11)*2^-53 - 1
    M = (1 << 64) - 1

    def sm(v):
        z = (v + 0x9E3779B97F4A7C15) & M
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return z ^ (z >> 31)
    for r in (0, 1, 31_415_926, dims[0] - 1):
        cols, vals = [], []
        for l in range(32):
            u = sm(42 ^ (r * 32 + l))
            cols.append((u * dims[1]) >> 64)
            vals.append(2.0 * ((sm(u) >> 11) * 2.0 ** -53) - 1.0)
        xs = x[torch.tensor(cols, device="cuda")].cpu().numpy()
        acc = 0.0
        for v, xv in zip(vals, xs):
            acc = acc + v * xv
        assert ye[r].item() == 0.0 + acc
