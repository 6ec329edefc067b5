"""GPU parity of the ELL path through the C ABI (include/ellspmv_cuda.h),
against the CPU oracle and the golden vectors of the unmodified reference.

Bar: ELL arrays bit-exact; y bit-exact in the default mode (thread-per-row,
mul-then-add); tolerance modes (FMA, sub-warp reduction) within the
dot-product bound  |dy| <= (K+2) * 2^-53 * sum_l |a_il * x_l|  per row."""
import numpy as np
import pytest

import ellspmv_b200 as E
from conftest import GOLDEN_CASES, bits_equal, load_golden, unhex

pytestmark = pytest.mark.gpu

ALL_R = [1, 2, 4]


def rand_ell(rng, nr, nc, K, dt, pad_frac=0.3):
    """Row-major ELL arrays with the reference's padding rule on a random tail of each row."""
    ec = rng.integers(0, nc, (nr, K)).astype(dt)
    ea = rng.standard_normal((nr, K))
    fill = rng.integers(0, K + 1, nr) if pad_frac > 0 else np.full(nr, K)
    for i in range(nr):
        ec[i, fill[i]:] = min(i, nc - 1)
        ea[i, fill[i]:] = 0.0
    return np.ascontiguousarray(ec.reshape(-1)), np.ascontiguousarray(ea.reshape(-1))


def tol_bound(K, ec, ea, x, nr):
    absprod = np.abs(ea.reshape(nr, K) * x[ec.reshape(nr, K)]).sum(axis=1) if K > 0 else np.zeros(nr)
    return (K + 2) * 2.0 ** -53 * absprod


@pytest.mark.parametrize("name", GOLDEN_CASES)
@pytest.mark.parametrize("bits", [32, 64])
def test_golden_reference_vectors(lib, name, bits):
    g = load_golden(name)
    e = g[f"idx{bits}"]
    dt = np.int32 if bits == 32 else np.int64
    ec, ea = np.array(e["ellcolidx"], dtype=dt), unhex(e["ella"])
    nr, nc, K = g["num_rows"], g["num_columns"], e["rowsize"]
    for R in ALL_R:
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, E.rows_per_thread(R))
        if name == "grid5":
            # the structured-grid fixture: most rows take their indices from an offset pattern,
            # and the result below is still the unmodified reference's, bit for bit
            assert A.info().pattern_rows == expected_pattern_rows(ec, K, nr, R) > 0, (R, A.info().pattern_rows)
        c2, a2 = A.download()
        assert np.array_equal(c2, ec) and bits_equal(a2, ea)
        y = unhex(g["y0"])
        A.spmv(y, unhex(g["x"]), 1, E.ACCUMULATE)
        assert bits_equal(y, unhex(e["y_ell"])), (name, bits, R)
        y = unhex(g["y0"])
        A.spmv(y, unhex(g["x"]), 3, E.ACCUMULATE)
        assert bits_equal(y, unhex(e["y_ell_repeat3"]))
        A.free()


def test_reference_shaped_operator(lib):
    g = load_golden("test_mtx")
    e = g["idx32"]
    ec, ea = np.array(e["ellcolidx"], dtype=np.int32), unhex(e["ella"])
    y, x = np.zeros(4), np.ones(5)
    assert E.ellgemv(4, y, 5, x, 20, 5, ec, ea) == 0
    assert y.tolist() == [3, 1, 3, 6]
    assert E.ellgemv(4, y, 5, x, 20, 5, ec, ea) == 0      # accumulates, like the reference
    assert y.tolist() == [6, 2, 6, 12]
    assert E.ellgemv(4, y, 5, x, 19, 5, ec, ea) != 0      # ellsize != N*K


SHAPES = [(1, 1, 1), (3, 7, 2), (127, 127, 5), (128, 64, 5), (129, 300, 7), (511, 511, 27), (512, 512, 32),
          (513, 100, 33), (1000, 1000, 1), (2049, 777, 16), (5000, 5000, 5), (4097, 4097, 27), (300, 5000, 64)]


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("bits", [32, 64])
def test_bit_exact_vs_oracle(lib, oracle, shape, bits):
    nr, nc, K = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr * 131 + nc * 7 + K + bits)
    ec, ea = rand_ell(rng, nr, nc, K, dt)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want = y0.copy()
    oracle.ellgemv(nr, want, x, K, ec, ea)
    want0 = np.zeros(nr)
    oracle.ellgemv(nr, want0, x, K, ec, ea)
    for R in ALL_R:
        for extra in (0, E.WIDE_INDEX, E.L2_PERSIST_X):
            A = E.EllMatrix.upload(nr, nc, K, ec, ea, E.rows_per_thread(R) | extra)
            info = A.info()
            assert info.rows_per_thread == R and info.slice_rows == 128 * R
            assert info.dev_idx_bits == (64 if (bits == 64 and extra == E.WIDE_INDEX) else 32)
            assert info.min_col == ec.min() and info.max_col == ec.max()
            c2, a2 = A.download()
            assert c2.dtype == dt and np.array_equal(c2, ec) and bits_equal(a2, ea)
            y = y0.copy()
            A.spmv(y, x, 1, E.ACCUMULATE)
            assert bits_equal(y, want), (shape, bits, R, extra)
            y = rng.standard_normal(nr)                    # garbage that OVERWRITE must ignore
            A.spmv(y, x, 1, E.OVERWRITE)
            assert bits_equal(y, want0), (shape, bits, R, extra)
            A.free()


@pytest.mark.parametrize("flags", [E.FMA, E.KERNEL_WARP, E.KERNEL_WARP | E.FMA])
@pytest.mark.parametrize("K", [3, 5, 16, 27, 32, 40, 100, 250])
def test_tolerance_modes(lib, oracle, flags, K):
    nr, nc = (3000, 2500) if K < 100 else (700, 2500)
    rng = np.random.default_rng(K + flags)
    ec, ea = rand_ell(rng, nr, nc, K, np.int32)
    x = rng.standard_normal(nc)
    want = np.zeros(nr)
    oracle.ellgemv(nr, want, x, K, ec, ea)
    bound = tol_bound(K, ec, ea, x, nr)
    for R in ALL_R:
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, flags | E.rows_per_thread(R))
        y = np.zeros(nr)
        A.spmv(y, x, 1, E.ACCUMULATE)
        A.free()
        assert np.all(np.abs(y - want) <= bound + 1e-300), (flags, K, R, np.max(np.abs(y - want)))


def test_edge_cases(lib, oracle):
    # no rows
    A = E.EllMatrix.upload(0, 5, 3, np.zeros(0, dtype=np.int32), np.zeros(0))
    A.spmv(np.zeros(0), np.ones(5), 2, E.ACCUMULATE)
    A.free()
    # K = 0: y unchanged under ACCUMULATE, zero under OVERWRITE
    A = E.EllMatrix.upload(7, 5, 0, np.zeros(0, dtype=np.int64), np.zeros(0))
    y = np.arange(7.0)
    A.spmv(y, np.ones(5), 2, E.ACCUMULATE)
    assert np.array_equal(y, np.arange(7.0))
    A.spmv(y, np.ones(5), 1, E.OVERWRITE)
    assert np.array_equal(y, np.zeros(7))
    A.free()
    # signed zeros: 0 + (-0) = +0 exactly like the CPU loop
    ec = np.zeros(4, dtype=np.int32)
    ea = np.array([-1.0, 0.0, 0.0, 0.0])
    x = np.array([0.0])
    want = np.array([-0.0])
    oracle.ellgemv(1, want, x, 4, ec, ea)
    A = E.EllMatrix.upload(1, 1, 4, ec, ea)
    y = np.array([-0.0])
    A.spmv(y, x, 1, E.ACCUMULATE)
    assert bits_equal(y, want)
    A.free()
    # inf / nan propagate like on the CPU (0 * inf = nan in a padded slot)
    ec = np.array([0, 1, 1, 0], dtype=np.int32)
    ea = np.array([1.0, 0.0, 2.0, 0.0])
    x = np.array([1.0, np.inf])
    want = np.zeros(2)
    oracle.ellgemv(2, want, x, 2, ec, ea)
    A = E.EllMatrix.upload(2, 2, 2, ec, ea)
    y = np.zeros(2)
    A.spmv(y, x, 1, E.ACCUMULATE)
    assert bits_equal(y, want) and np.isnan(y[0]) and np.isinf(y[1])
    A.free()
    # out-of-range column index is rejected at upload
    with pytest.raises(E.EllspmvCudaError):
        E.EllMatrix.upload(1, 3, 1, np.array([3], dtype=np.int32), np.ones(1))


@pytest.mark.parametrize("bits", [32, 64])
def test_iterate_mode(lib, oracle, bits):
    dt = np.int32 if bits == 32 else np.int64
    K, ncols, ec, ea, _ = oracle.gen_ell("stencil27", (12, 10, 11), (0.5, -1.0 / 52), bits=bits)
    n = ncols
    x0 = np.random.default_rng(5).uniform(-1, 1, n)
    want = oracle.ell_iterate(n, x0, 7, K, ec, ea)
    A = E.EllMatrix.upload(n, n, K, ec, ea)
    y = np.zeros(n)
    A.spmv(y, x0, 7, E.ITERATE)
    assert bits_equal(y, want)
    # even count too (result lives in the other buffer)
    want = oracle.ell_iterate(n, x0, 4, K, ec, ea)
    A.spmv(y, x0, 4, E.ITERATE)
    assert bits_equal(y, want)
    A.free()
    B = E.EllMatrix.upload(3, 4, 1, np.zeros(3, dtype=dt), np.ones(3))
    with pytest.raises(E.EllspmvCudaError):
        B.spmv(np.zeros(3), np.ones(4), 1, E.ITERATE)    # not square
    B.free()


def test_device_vectors_and_shards(lib, oracle):
    """Row shards with global column indices give the same bits as the whole
    matrix (row sharding leaves each row's summation order untouched)."""
    import torch
    nr, nc, K = 5003, 5003, 9
    rng = np.random.default_rng(11)
    ec, ea = rand_ell(rng, nr, nc, K, np.int32)
    x = rng.standard_normal(nc)
    want = np.zeros(nr)
    oracle.ellgemv(nr, want, x, K, ec, ea)
    xd = torch.from_numpy(x).cuda()
    yd = torch.full((nr,), 123.0, dtype=torch.float64, device="cuda")
    cuts = [0, 1000, 1003, 3500, nr]
    for lo, hi in zip(cuts[:-1], cuts[1:]):
        S = E.EllMatrix.upload(hi - lo, nc, K, ec[lo * K:hi * K], ea[lo * K:hi * K], 0,
                               global_rows=nr, row_begin=lo, device=0)
        i = S.info()
        assert (i.row_begin, i.num_rows, i.global_rows) == (lo, hi - lo, nr)
        S.spmv_device(yd[lo:hi], xd, E.OVERWRITE, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        S.free()
    assert bits_equal(yd.cpu().numpy(), want)


def test_push_to_peer_vectors(lib, oracle):
    """The fused exchange on one GPU: 'peers' are plain device vectors."""
    import torch
    nr = nc = 4099
    K = 6
    rng = np.random.default_rng(12)
    ec, ea = rand_ell(rng, nr, nc, K, np.int64)
    x = rng.standard_normal(nc)
    want = np.zeros(nr)
    oracle.ellgemv(nr, want, x, K, ec, ea)
    xd = torch.from_numpy(x).cuda()
    lo, hi = 1001, 3001      # deliberately not a multiple of 4
    S = E.EllMatrix.upload(hi - lo, nc, K, ec[lo * K:hi * K], ea[lo * K:hi * K], 0,
                           global_rows=nr, row_begin=lo, device=0)
    own = torch.zeros(nr, dtype=torch.float64, device="cuda")
    p0 = torch.zeros(nr, dtype=torch.float64, device="cuda")
    p1 = torch.zeros(nr, dtype=torch.float64, device="cuda")
    S.spmv_push(own[lo:hi], xd, E.OVERWRITE, [p0.data_ptr(), p1.data_ptr()], [0, 2000], [nr, 2500],
                torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    S.free()
    assert bits_equal(own[lo:hi].cpu().numpy(), want[lo:hi])
    assert bits_equal(p0[lo:hi].cpu().numpy(), want[lo:hi]) and p0[:lo].abs().sum() == 0 and p0[hi:].abs().sum() == 0
    assert bits_equal(p1[2000:2500].cpu().numpy(), want[2000:2500])
    assert p1[:2000].abs().sum() == 0 and p1[2500:].abs().sum() == 0


@pytest.mark.parametrize("fused", [0, 1])
def test_exchange_on_one_gpu(lib, oracle, fused):
    """ellspmv_cuda_spmv_exchange with two shards of one matrix on ONE device, each playing a
    rank: SpMV + push + the step hand-shake (one-warp kernel after the SpMV kernel, or inside it
    with FUSED_SYNC), several steps of x <- A*x, bits of the oracle's iteration.  (Both 'ranks' run
    on the same GPU on two streams, so whoever waits is really waiting for the other shard.)"""
    import torch
    for name, dims, vals, bits in [("laplace2d", (700, 97), (0.25, -0.125), 32),
                                   ("stencil27", (30, 9, 11), (0.5, -1.0 / 52), 64)]:
        K, ncols, ec, ea, _ = oracle.gen_ell(name, dims, vals, bits=bits)
        rows = len(ea) // K
        x0 = np.random.default_rng(5).uniform(-1, 1, rows)
        steps = 7
        want = oracle.ell_iterate(rows, x0, steps, K, ec, ea)
        cut = rows // 2 + 3                                   # not a multiple of 16
        parts = [(0, cut), (cut, rows)]
        S = [E.EllMatrix.upload(b - a, ncols, K, ec[a * K:b * K], ea[a * K:b * K], E.FUSED_SYNC if fused else 0,
                                global_rows=rows, row_begin=a, device=0) for a, b in parts]
        needs = [(S[r].info().min_col, S[r].info().max_col + 1) for r in range(2)]
        # each rank owns two full-length vectors and a flag array
        xb = [[torch.zeros(rows, dtype=torch.float64, device="cuda") for _ in range(2)] for _ in range(2)]
        flags = [torch.zeros(32, dtype=torch.int64, device="cuda") for _ in range(2)]
        for r in range(2):
            xb[r][0].copy_(torch.from_numpy(x0))
        torch.cuda.synchronize()
        streams = [torch.cuda.Stream(), torch.cuda.Stream()]
        cur = 0
        for k in range(steps):
            for r in range(2):
                o = 1 - r
                a, b = parts[r]
                lo, hi = max(a, needs[o][0]), min(b, needs[o][1])
                S[r].spmv_exchange(xb[r][1 - cur][a:b], xb[r][cur], E.OVERWRITE, [xb[o][1 - cur].data_ptr()], [lo], [hi],
                                   r, [o], [flags[o].data_ptr()], flags[r].data_ptr(), k + 1, streams[r].cuda_stream)
            cur = 1 - cur
        torch.cuda.synchronize()
        got = torch.cat([xb[0][cur][:cut], xb[1][cur][cut:]]).cpu().numpy()
        assert bits_equal(got, want), name
        assert int(flags[0][1]) == steps and int(flags[1][0]) == steps and int(flags[0][16]) == 0
        for m in S:
            m.free()


def test_host_call_uploads_only_the_referenced_x_range(lib, oracle):
    """A row shard's host-vector call reads x only on the column range the shard references
    (its rows and the halo): poison everything else."""
    K, ncols, ec, ea, _ = oracle.gen_ell("laplace2d", (1500, 1001), (4.0, -1.0), bits=32)
    rows = len(ea) // K
    rng = np.random.default_rng(6)
    x = rng.standard_normal(ncols)
    a, b = rows // 3, 2 * rows // 3 + 5
    want = np.zeros(b - a)
    oracle.ellgemv(b - a, want, x, K, ec[a * K:b * K], ea[a * K:b * K])
    S = E.EllMatrix.upload(b - a, ncols, K, ec[a * K:b * K], ea[a * K:b * K], 0, global_rows=rows, row_begin=a, device=0)
    i = S.info()
    xp = np.full(ncols, np.nan)
    xp[i.min_col:i.max_col + 1] = x[i.min_col:i.max_col + 1]
    assert i.max_col - i.min_col + 1 < ncols // 2
    for repeat in (1, 2):          # the pipelined path (>= 2^20 rows would take it) and the plain one
        y = np.zeros(b - a)
        S.spmv(y, xp, repeat, E.OVERWRITE)
        assert bits_equal(y, want)
    S.free()


def test_negative_column_indices_are_rejected(lib):
    """All-negative indices used to slip through the range check (the max accumulator starts at -1)."""
    for dt in (np.int32, np.int64):
        ec = np.full(6, -3, dtype=dt)
        with pytest.raises(E.EllspmvCudaError) as ei:
            E.EllMatrix.upload(3, 10, 2, ec, np.ones(6))
        assert ei.value.errno == 22
        ec = np.array([0, 1, -1, 2, 3, 4], dtype=dt)
        with pytest.raises(E.EllspmvCudaError):
            E.EllMatrix.upload(3, 10, 2, ec, np.ones(6))
        ec = np.array([0, 1, 10, 2, 3, 4], dtype=dt)
        with pytest.raises(E.EllspmvCudaError):
            E.EllMatrix.upload(3, 10, 2, ec, np.ones(6))


@pytest.mark.parametrize("mode", [E.ACCUMULATE, E.OVERWRITE])
def test_pipelined_host_call(lib, oracle, mode):
    """>= 2^20 rows and repeat == 1 take the chunked, copy-overlapped path of
    ellspmv_cuda_spmv; pinned and pageable host vectors, ragged last chunk."""
    import torch
    K, ncols, ec, ea, _ = oracle.gen_ell("laplace2d", (1500, 1001), (4.0, -1.0), bits=32)
    rows = len(ea) // K
    rng = np.random.default_rng(4)
    x = rng.standard_normal(ncols)
    y0 = rng.standard_normal(rows)
    want = y0.copy() if mode == E.ACCUMULATE else np.zeros(rows)
    oracle.ellgemv(rows, want, x, K, ec, ea)
    A = E.EllMatrix.upload(rows, ncols, K, ec, ea)
    y = y0.copy()
    secs = A.spmv(y, x, 1, mode)
    assert bits_equal(y, want) and secs[0] > 0
    xp = torch.from_numpy(x).pin_memory()
    yp = torch.from_numpy(y0.copy()).pin_memory()
    A.spmv(yp.numpy(), xp.numpy(), 1, mode)
    assert bits_equal(yp.numpy(), want)
    A.free()
    # a row shard (>= 2^20 rows) through the same path: only its x range is uploaded
    a, b = 100_003, rows - 77
    S = E.EllMatrix.upload(b - a, ncols, K, ec[a * K:b * K], ea[a * K:b * K], 0, global_rows=rows, row_begin=a, device=0)
    i = S.info()
    xq = np.full(ncols, np.nan)
    xq[i.min_col:i.max_col + 1] = x[i.min_col:i.max_col + 1]
    y = y0[a:b].copy()
    S.spmv(y, xq, 1, mode)
    assert bits_equal(y, want[a:b])
    S.free()


@pytest.mark.parametrize("shape", [(1, 1, 1), (127, 300, 5), (1000, 1000, 7), (5003, 5003, 27), (40000, 40000, 5), (3000, 3000, 32)])
@pytest.mark.parametrize("bits", [32, 64])
def test_bulk_async_variant_is_bit_exact(lib, oracle, shape, bits):
    """ELLSPMV_CUDA_VARIANT = 1: persistent CTAs fed by cp.async.bulk + mbarrier
    (TMA) instead of per-thread vector loads; same bits as the oracle."""
    nr, nc, K = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + nc + K + bits)
    ec, ea = rand_ell(rng, nr, nc, K, dt)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want = y0.copy()
    oracle.ellgemv(nr, want, x, K, ec, ea)
    want0 = np.zeros(nr)
    oracle.ellgemv(nr, want0, x, K, ec, ea)
    for R in ALL_R:
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, E.rows_per_thread(R) | E.variant(1))
        y = y0.copy()
        A.spmv(y, x, 2, E.ACCUMULATE)
        want2 = want.copy()
        oracle.ellgemv(nr, want2, x, K, ec, ea)
        assert bits_equal(y, want2), (shape, bits, R)
        A.spmv(y, x, 1, E.OVERWRITE)
        assert bits_equal(y, want0), (shape, bits, R)
        A.free()


@pytest.mark.parametrize("shape", [(1, 40, 3), (1000, 1000, 7), (5003, 9001, 27), (300, 70000, 32), (4097, 4097, 64)])
@pytest.mark.parametrize("bits", [32, 64])
def test_column_blocked_mode(lib, oracle, shape, bits, monkeypatch):
    """ELLSPMV_CUDA_COLUMN_BLOCKED (tolerance mode): entries binned by column
    block, y += A_b*x block after block.  Forced to many small blocks here."""
    nr, nc, K = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + nc + K + bits)
    ec, ea = rand_ell(rng, nr, nc, K, dt)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want0 = np.zeros(nr)
    oracle.ellgemv(nr, want0, x, K, ec, ea)
    absprod = np.abs(ea.reshape(nr, K) * x[ec.reshape(nr, K)]).sum(axis=1)
    bound = (K + 2 + 64) * 2.0 ** -53 * absprod            # up to 64 extra additions of block partial sums
    for block_bytes in (256, 8 * 1024, 1 << 30):
        monkeypatch.setenv("ELLSPMV_CUDA_BLOCK_BYTES", str(block_bytes))
        for flags in (E.COLUMN_BLOCKED, E.COLUMN_BLOCKED | E.FMA):
            A = E.EllMatrix.upload(nr, nc, K, ec, ea, flags)
            c2, a2 = A.download()                       # the regular layout is still there, bit for bit
            assert np.array_equal(c2, ec) and bits_equal(a2, ea)
            y = y0.copy()
            A.spmv(y, x, 1, E.ACCUMULATE)
            assert np.all(np.abs(y - (y0 + want0)) <= bound + 4e-16 * np.abs(y0) + 1e-300), (shape, bits, block_bytes, flags)
            y = rng.standard_normal(nr)
            A.spmv(y, x, 1, E.OVERWRITE)
            assert np.all(np.abs(y - want0) <= bound + 1e-300)
            if block_bytes == 1 << 30 and flags == E.COLUMN_BLOCKED:
                assert bits_equal(y, want0)             # one block: falls back to the bit-exact kernel
            A.free()


@pytest.mark.parametrize("shape", [(1, 40, 3), (1000, 1000, 7), (5003, 9001, 27), (300, 70000, 32), (4097, 4097, 64),
                                   (777, 5000, 150)])
@pytest.mark.parametrize("bits", [32, 64])
def test_staged_gather_is_bit_exact(lib, oracle, shape, bits, monkeypatch):
    """ELLSPMV_CUDA_STAGED_GATHER: column blocks with the gather staged through
    device memory keep the reference's summation order, so the result is the
    oracle's bit for bit -- with many small blocks, with one block (falls back to
    the plain kernel), with R = 2/4 slices, wide indices and a separate diagonal."""
    nr, nc, K = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + nc + K + bits + 1)
    ec, ea = rand_ell(rng, nr, nc, K, dt)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want0 = np.zeros(nr)
    oracle.ellgemv(nr, want0, x, K, ec, ea)
    want1 = y0.copy()
    oracle.ellgemv(nr, want1, x, K, ec, ea)
    want2 = want1.copy()
    oracle.ellgemv(nr, want2, x, K, ec, ea)
    for block_bytes in (256, 8 * 1024, 1 << 30):
        monkeypatch.setenv("ELLSPMV_CUDA_BLOCK_BYTES", str(block_bytes))
        for flags in (E.STAGED_GATHER, E.STAGED_GATHER | E.rows_per_thread(2), E.STAGED_GATHER | E.rows_per_thread(4),
                      E.STAGED_GATHER | E.WIDE_INDEX, E.STAGED_GATHER | E.COLUMN_BLOCKED):
            A = E.EllMatrix.upload(nr, nc, K, ec, ea, flags)
            c2, a2 = A.download()
            assert np.array_equal(c2, ec) and bits_equal(a2, ea)
            y = y0.copy()
            A.spmv(y, x, 2, E.ACCUMULATE)
            assert bits_equal(y, want2), (shape, bits, block_bytes, flags)
            y = rng.standard_normal(nr)
            A.spmv(y, x, 1, E.OVERWRITE)
            assert bits_equal(y, want0), (shape, bits, block_bytes, flags)
            A.free()


@pytest.mark.parametrize("shape", [(1, 5000, 5000), (3, 100, 70), (200, 3000, 257), (1000, 1000, 64), (40, 9000, 4097),
                                   (33, 400, 1024), (5000, 5000, 96)])
@pytest.mark.parametrize("bits", [32, 64])
def test_long_row_kernel_is_bit_exact(lib, oracle, shape, bits):
    """Few, long rows: KERNEL_AUTO takes the CTA-per-row-group kernel on the row-major layout
    (ell_longrow.cu); products parked in shared memory, one lane adds them in slot order -> the
    oracle's bits, for every rows-per-CTA choice, accumulate and overwrite, both diagonal orders."""
    nr, nc, K = shape
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(nr + K + bits)
    ec, ea = rand_ell(rng, nr, nc, K, dt)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want = y0.copy()
    oracle.ellgemv(nr, want, x, K, ec, ea)
    want0 = np.zeros(nr)
    oracle.ellgemv(nr, want0, x, K, ec, ea)
    for flags in (0, E.KERNEL_LONGROW):
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, flags)
        i = A.info()
        assert i.kernel == E.KERNEL_LONGROW and i.slice_rows == 1 and i.long_rows == K, (flags, i.kernel)
        c2, a2 = A.download()
        assert np.array_equal(c2, ec) and bits_equal(a2, ea)
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        assert bits_equal(y, want)
        A.spmv(y, x, 1, E.OVERWRITE)
        assert bits_equal(y, want0)
        if nr <= nc:
            ad = rng.standard_normal(nr)
            for order in (0, 1):
                w = y0.copy()
                oracle.ellgemvsd(nr, w, x, K, ec, ea, ad, order)
                A.set_diagonal(ad, order)
                y = y0.copy()
                A.spmv(y, x, 1, E.ACCUMULATE)
                assert bits_equal(y, w), order
        A.free()
    # the thread-per-row kernel on the same matrix: same bits (the switch is invisible)
    B = E.EllMatrix.upload(nr, nc, K, ec, ea, E.KERNEL_THREAD)
    assert B.info().kernel == E.KERNEL_THREAD
    y = y0.copy()
    B.spmv(y, x, 1, E.ACCUMULATE)
    B.free()
    assert bits_equal(y, want)
    # special values travel the same way
    xs = x.copy()
    xs[rng.integers(0, nc, 3)] = np.inf
    w = np.zeros(nr)
    oracle.ellgemv(nr, w, xs, K, ec, ea)
    C2 = E.EllMatrix.upload(nr, nc, K, ec, ea)
    y = np.zeros(nr)
    C2.spmv(y, xs, 1, E.ACCUMULATE)
    C2.free()
    assert np.array_equal(np.isnan(y), np.isnan(w))
    ok = ~np.isnan(w)
    assert bits_equal(y[ok], w[ok])


@pytest.mark.parametrize("variant", [1, 2, 3, 0])
@pytest.mark.parametrize("rshift", [0, 1, 2, 3, 4, 5])
def test_long_row_kernel_every_rows_per_cta(lib, oracle, monkeypatch, variant, rshift):
    """The forms of the long-row kernel (1: loader warps + a summing warp over an mbarrier ring, the
    default; 3: the same without 16-byte copies; 2: the first ring build; 0: the lock-step form) at
    every rows-per-CTA choice: ragged last CTA, K not a multiple of
    the stage, one tile, many tiles, 64-bit indices, per-row lengths through the CSR view."""
    monkeypatch.setenv("ELLSPMV_CUDA_LONGROW_VARIANT", str(variant))
    monkeypatch.setenv("ELLSPMV_CUDA_LONGROW_RSHIFT", str(rshift))
    # K a multiple of 4 (of 2 with 64-bit indices kept wide): full stages go up in 16-byte copies
    for (nr, nc, K), dt in (((37, 5000, 3001), np.int32), ((70, 300, 64), np.int64), ((5, 20000, 10241), np.int32),
                            ((1, 7, 1), np.int32), ((129, 4000, 1024), np.int64), ((64, 3000, 2052), np.int32),
                            ((96, 900, 1030), np.int32)):
        rng = np.random.default_rng(nr * 7 + K + rshift)
        ec, ea = rand_ell(rng, nr, nc, K, dt)
        x = rng.standard_normal(nc)
        y0 = rng.standard_normal(nr)
        want = y0.copy()
        oracle.ellgemv(nr, want, x, K, ec, ea)
        # 64-bit indices stay 64-bit on the device here (the kernel's int64 instantiation)
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, E.KERNEL_LONGROW | (E.WIDE_INDEX if dt == np.int64 else 0))
        assert A.info().kernel == E.KERNEL_LONGROW and A.info().dev_idx_bits == ec.dtype.itemsize * 8
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        A.free()
        assert bits_equal(y, want), (nr, nc, K, variant, rshift)
    # CSR rows of different lengths (few, long): the view hands the kernel per-row lengths
    rng = np.random.default_rng(rshift)
    nr, nc = 50, 6000
    lens = rng.integers(1500, 1900, nr)
    rowptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    cols = rng.integers(0, nc, rowptr[-1]).astype(np.int32)
    vals = rng.standard_normal(rowptr[-1])
    x = rng.standard_normal(nc)
    x[rng.integers(0, nc, 2)] = np.inf
    want = np.zeros(nr)
    oracle.csrgemv(nr, want, x, rowptr, cols, vals)
    C = E.CsrMatrix.upload(nr, nc, rowptr, cols, vals)
    y = np.zeros(nr)
    C.spmv(y, x, 1, E.ACCUMULATE)
    view = C.info().ell_view
    C.free()
    assert view == 1                      # the sliced-ELL view with per-row lengths; 50 rows -> the long-row kernel
    assert np.array_equal(np.isnan(y), np.isnan(want))
    ok = ~np.isnan(want)
    assert bits_equal(y[ok], want[ok]), (variant, rshift)


def test_long_row_kernel_in_a_shard_and_iterate(lib, oracle):
    import torch
    nr = nc = 900
    K = 130
    rng = np.random.default_rng(3)
    ec, ea = rand_ell(rng, nr, nc, K, np.int32)
    ea *= 0.01
    x = rng.standard_normal(nc)
    want = oracle.ell_iterate(nr, x, 4, K, ec, ea)
    A = E.EllMatrix.upload(nr, nc, K, ec, ea)
    assert A.info().kernel == E.KERNEL_LONGROW
    y = np.zeros(nr)
    A.spmv(y, x, 4, E.ITERATE)
    A.free()
    assert bits_equal(y, want)
    lo, hi = 123, 777
    S = E.EllMatrix.upload(hi - lo, nc, K, ec[lo * K:hi * K], ea[lo * K:hi * K], global_rows=nr, row_begin=lo, device=0)
    w = np.zeros(nr)
    oracle.ellgemv(nr, w, x, K, ec, ea)
    xd = torch.from_numpy(x).cuda()
    own = torch.zeros(nr, dtype=torch.float64, device="cuda")
    peer = torch.zeros(nr, dtype=torch.float64, device="cuda")
    S.spmv_push(own[lo:hi], xd, E.OVERWRITE, [peer.data_ptr()], [200], [300], torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    S.free()
    assert bits_equal(own[lo:hi].cpu().numpy(), w[lo:hi]) and bits_equal(peer[200:300].cpu().numpy(), w[200:300])
    assert peer[:200].abs().sum() == 0 and peer[300:].abs().sum() == 0


@pytest.mark.parametrize("bits", [32, 64])
def test_skip_padding_runs_from_the_sell_copy(lib, oracle, bits):
    """ELLSPMV_CUDA_SKIP_PADDING: rows sorted by their length without trailing padding, a width per
    slice; only the slots that count are streamed.  Same bits as the reference for finite x; download
    and the other entry points still see the padded ELL arrays."""
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(40 + bits)
    nr, nc, K = 9000, 9500, 40
    ec = np.empty((nr, K), dtype=dt)
    ea = np.zeros((nr, K))
    lens = np.minimum((2 / rng.random(nr) ** 0.8).astype(np.int64), K)      # ragged: most rows short, a few full
    lens[rng.random(nr) < 0.03] = 0
    for i in range(nr):
        n = int(lens[i])
        ec[i, :n] = rng.integers(0, nc, n)
        ea[i, :n] = rng.standard_normal(n)
        ec[i, n:] = min(i, nc - 1)                                       # the reference's padding rule
    ea[5, 1] = 0.0                                                      # a stored zero INSIDE a row is not padding
    ec, ea = ec.reshape(-1), ea.reshape(-1)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want = y0.copy()
    oracle.ellgemv(nr, want, x, K, ec, ea)
    A = E.EllMatrix.upload(nr, nc, K, ec, ea, E.SKIP_PADDING)
    i = A.info()
    assert 0 < i.sell_slots < nr * K // 3, i.sell_slots
    y = y0.copy()
    A.spmv(y, x, 1, E.ACCUMULATE)
    assert bits_equal(y, want)
    y = np.full(nr, 9.0)
    A.spmv(y, x, 1, E.OVERWRITE)
    w0 = np.zeros(nr)
    oracle.ellgemv(nr, w0, x, K, ec, ea)
    assert bits_equal(y, w0)
    c2, a2 = A.download()
    assert np.array_equal(c2, ec) and bits_equal(a2, ea)
    for order in (0, 1):                      # order 1 falls back to the padded layout: same bits either way
        ad = rng.standard_normal(nr)
        w = y0.copy()
        oracle.ellgemvsd(nr, w, x, K, ec, ea, ad, order)
        A.set_diagonal(ad, order)
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        assert bits_equal(y, w), order
    A.free()


def test_auto_tries_the_staged_gather_on_scattered_matrices(lib, oracle, monkeypatch):
    """KERNEL_AUTO: x larger than the threshold + no offset patterns + scattered gathers -> the staged
    copy is built and timed against the direct gather at upload; whichever is kept, the bits are the
    oracle's.  A banded matrix (gathers share lines) and NO_STAGED_GATHER skip the trial."""
    monkeypatch.setenv("ELLSPMV_CUDA_AUTO_STAGED_MIN_X_BYTES", "4096")
    monkeypatch.setenv("ELLSPMV_CUDA_BLOCK_BYTES", "65536")
    rng = np.random.default_rng(21)
    nr, nc, K = 6000, 70000, 32
    ec, ea = rand_ell(rng, nr, nc, K, np.int32, pad_frac=0.0)
    x = rng.standard_normal(nc)
    want = np.zeros(nr)
    oracle.ellgemv(nr, want, x, K, ec, ea)
    A = E.EllMatrix.upload(nr, nc, K, ec, ea)
    i = A.info()
    assert i.tune_ms[0] > 0 and i.tune_ms[1] > 0 and i.staged in (0, 2)
    assert i.launches_per_spmv == (1 if i.staged == 0 else 10)     # 70000 * 8 B / 64 KB -> 9 column blocks + the sum
    y = np.zeros(nr)
    A.spmv(y, x, 1, E.ACCUMULATE)
    A.free()
    assert bits_equal(y, want)
    B = E.EllMatrix.upload(nr, nc, K, ec, ea, E.NO_STAGED_GATHER)
    assert B.info().tune_ms[0] == 0 and B.info().staged == 0
    B.free()
    bc, ba = banded_ell(nr, nr, [-3, -1, 0, 1, 2], np.int32, rng)
    Cm = E.EllMatrix.upload(nr, nr, 5, bc, ba, E.NO_PATTERN)
    assert Cm.info().tune_ms[0] == 0 and Cm.info().staged == 0
    Cm.free()
    D = E.EllMatrix.upload(nr, nc, K, ec, ea, E.STAGED_GATHER)
    assert D.info().staged == 1
    D.free()


def test_staged_gather_tolerance_and_special_values(lib, oracle, monkeypatch):
    """FMA on the staged path stays inside the dot-product bound; inf/NaN in x
    propagate exactly like in the reference loop (no stored zero is dropped)."""
    nr, nc, K = 3000, 4000, 9
    rng = np.random.default_rng(99)
    ec, ea = rand_ell(rng, nr, nc, K, np.int32)
    ea[rng.random(ea.shape) < 0.2] = 0.0                # explicit zeros / padding-like entries
    x = rng.standard_normal(nc)
    monkeypatch.setenv("ELLSPMV_CUDA_BLOCK_BYTES", "4096")
    want = np.zeros(nr)
    oracle.ellgemv(nr, want, x, K, ec, ea)
    A = E.EllMatrix.upload(nr, nc, K, ec, ea, E.STAGED_GATHER | E.FMA)
    y = np.zeros(nr)
    A.spmv(y, x, 1, E.OVERWRITE)
    absprod = np.abs(ea.reshape(nr, K) * x[ec.reshape(nr, K)]).sum(axis=1)
    assert np.all(np.abs(y - want) <= (K + 2) * 2.0 ** -53 * absprod + 1e-300)
    A.free()
    # NaNs are only ever PRODUCED here (0*inf, inf-inf): a produced NaN has the same bits on
    # x86 and on the GPU, whereas an input NaN's payload is kept by SSE2 and canonicalised
    # by the GPU's DMUL/DADD (true of every kernel in this library)
    x[::7] = np.inf
    x[3::11] = -np.inf
    x[5::13] = -0.0
    want = np.zeros(nr)
    with np.errstate(all="ignore"):
        oracle.ellgemv(nr, want, x, K, ec, ea)
    assert np.isnan(want).any() and np.isinf(want).any()
    for flags in (0, E.STAGED_GATHER):
        A = E.EllMatrix.upload(nr, nc, K, ec, ea, flags)
        y = np.zeros(nr)
        A.spmv(y, x, 1, E.OVERWRITE)
        assert bits_equal(y, want), flags
        A.free()


def banded_ell(nr, nc, offsets, dt, rng):
    """ELL arrays of a matrix whose row i holds columns i + d for d in offsets (clipped rows get
    the reference's padding: column min(i, nc-1), value 0)."""
    K = len(offsets)
    ec = np.empty((nr, K), dtype=dt)
    ea = np.zeros((nr, K))
    for i in range(nr):
        cols = [i + d for d in offsets if 0 <= i + d < nc]
        ec[i, :len(cols)] = cols
        ea[i, :len(cols)] = rng.standard_normal(len(cols))
        ec[i, len(cols):] = min(i, nc - 1)
    return ec.reshape(-1), ea.reshape(-1)


@pytest.mark.parametrize("bits", [32, 64])
def test_offset_patterns_are_found_and_change_nothing(lib, oracle, bits):
    """Groups of 32*R rows (one warp) whose column indices are row + d[l] take them from the
    pattern dictionary instead of the index stream (pattern.cu).  Same bits with and without, for
    narrowed and wide indices; download() still returns the explicit arrays."""
    dt = np.int32 if bits == 32 else np.int64
    rng = np.random.default_rng(7 + bits)
    nr = nc = 5000
    ec, ea = banded_ell(nr, nc, (-70, -1, 0, 1, 70), dt, rng)
    x = rng.standard_normal(nc)
    y0 = rng.standard_normal(nr)
    want = y0.copy()
    oracle.ellgemv(nr, want, x, 5, ec, ea)
    for flags, expect_patterns in ((0, True), (E.WIDE_INDEX, True), (E.FMA | E.NO_PATTERN, False), (E.NO_PATTERN, False),
                                   (E.rows_per_thread(2), True), (E.rows_per_thread(4), True), (E.KERNEL_WARP, False)):
        A = E.EllMatrix.upload(nr, nc, 5, ec, ea, flags)
        rows = A.info().pattern_rows
        # interior groups: rows 70..4929 minus the groups straddling the two boundaries
        assert (rows >= 4700) if expect_patterns else (rows == 0), (flags, rows)
        c2, a2 = A.download()
        assert np.array_equal(c2, ec) and bits_equal(a2, ea)
        y = y0.copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        if flags & (E.FMA | E.KERNEL_WARP):
            assert np.allclose(y, want, rtol=1e-13, atol=1e-13)
        else:
            assert bits_equal(y, want), flags
        A.free()


def test_offset_patterns_adversarial(lib, oracle):
    """One deviating entry keeps its group on the explicit stream -- or, with ELLSPMV_CUDA_PATTERN_MASKS,
    only its own row; more distinct patterns than the dictionary holds leaves the rest explicit; a
    random matrix finds nothing.  Always the oracle's bits."""
    rng = np.random.default_rng(11)
    nr = nc = 4096
    # (a) a single entry of a single row differs from its group's pattern
    ec, ea = banded_ell(nr, nc, (-3, 0, 5), np.int32, rng)
    ec = ec.reshape(nr, 3)
    ec[1000, 2] = 17
    ec[2049, 0] = 2049            # same column as slot 1: a duplicate column inside the row
    ec = ec.reshape(-1)
    # (b) 40 stripes of rows, each with its own offsets: more than 16 patterns
    ec2 = np.empty((nr, 2), dtype=np.int32)
    for i in range(nr):
        stripe = i // 96
        ec2[i] = ((i + stripe) % nc, (i + 2 * stripe + 1) % nc)
    ea2 = rng.standard_normal(nr * 2)
    # (c) random
    ec3 = rng.integers(0, nc, nr * 4).astype(np.int32)
    ea3 = rng.standard_normal(nr * 4)
    x = rng.standard_normal(nc)
    # (a) with lane masks: the two damaged rows and the boundary rows stay explicit, their groups stay
    # patterned (the last group has 5 boundary rows, more than a mask takes: explicit as a whole)
    for K, cols, vals, lo, hi, mlo, mhi in ((3, ec, ea, 3900, 4096 - 64, 4096 - 48, 4096 - 34),
                                            (2, ec2.reshape(-1), ea2, 32, 4096 - 1024, 32, 4096 - 1024),
                                            (4, ec3, ea3, 0, 0, 0, 0)):
        want = np.zeros(nr)
        oracle.ellgemv(nr, want, x, K, cols, vals)
        for flags, a, b, me in ((0, lo, hi, 0), (E.PATTERN_MASKS, mlo, mhi, 4)):
            A = E.EllMatrix.upload(nr, nc, K, cols, vals, E.rows_per_thread(1) | flags)    # groups of 32 rows
            rows = A.info().pattern_rows
            assert a <= rows <= b, (K, rows, flags)
            assert rows == expected_pattern_rows(cols, K, nr, 1, max_explicit=me), (K, flags)
            y = np.zeros(nr)
            A.spmv(y, x, 1, E.OVERWRITE)
            assert bits_equal(y, want), (K, flags)
            A.free()


@pytest.mark.parametrize("bits", [32, 64])
def test_lane_patterns_on_small_grids(lib, oracle, bits):
    """Grids whose lines are shorter than a group: every group of 32 rows holds a boundary row, so
    whole-group ids find (almost) nothing and one id per thread covers the matrix.  The rows on
    the dictionary are the numpy restatement's, y is the oracle's bit for bit -- whole matrix, row
    shard, with NO_PATTERN_LANES and NO_PATTERN beside it."""
    rng = np.random.default_rng(77)
    for kind, dims, K in ((E.GEN_STENCIL27, (24, 20, 16), 27), (E.GEN_STENCIL27, (48, 9, 40), 27),
                          (E.GEN_LAPLACE2D, (96, 50), 5)):
        G = E.EllMatrix.generate(kind, dims, (26.0, -1.0), 42, bits, flags=E.NO_PATTERN)
        nr = int(np.prod(dims))
        ec, ea = G.download()
        G.free()
        ea = rng.standard_normal(ea.shape) * (ea != 0.0)       # variable coefficients, padding stays zero
        x = rng.standard_normal(nr)
        want = rng.standard_normal(nr)
        y0 = want.copy()
        oracle.ellgemv(nr, want, x, K, ec, ea)
        seen = {}
        for R in (1, 2):
            for flags, lanes in ((0, True), (E.NO_PATTERN_LANES, False)):
                A = E.EllMatrix.upload(nr, nr, K, ec, ea, E.rows_per_thread(R) | flags)
                rows = A.info().pattern_rows
                assert rows == expected_pattern_rows(ec, K, nr, R, lanes=lanes), (kind, dims, R, lanes, rows)
                seen[(R, lanes)] = rows
                y = y0.copy()
                A.spmv(y, x, 1, E.ACCUMULATE)
                assert bits_equal(y, want), (kind, dims, R, lanes)
                c2, a2 = A.download()
                assert np.array_equal(c2, ec) and bits_equal(a2, ea)
                A.free()
        assert seen[(1, True)] > seen[(1, False)] and seen[(1, True)] >= 0.85 * nr, (kind, dims, seen)
        # a row shard: local rows, global columns, an odd first row (y not vector-aligned)
        lo, hi = nr // 3 + 1, nr - nr // 5
        A = E.EllMatrix.upload(hi - lo, nr, K, ec[lo * K:hi * K], ea[lo * K:hi * K], E.rows_per_thread(1),
                               global_rows=nr, row_begin=lo)
        assert A.info().pattern_rows == expected_pattern_rows(ec[lo * K:hi * K], K, hi - lo, 1, row_begin=lo)
        y = y0[lo:hi].copy()
        A.spmv(y, x, 1, E.ACCUMULATE)
        assert bits_equal(y, want[lo:hi])
        A.free()


@pytest.mark.parametrize("bits", [32, 64])
def test_value_patterns_are_opt_in_and_bit_exact(lib, oracle, bits):
    """ELLSPMV_CUDA_VALUE_PATTERN: rows that share offsets AND coefficients take both from the
    dictionary.  Off by default (value_pattern_rows == 0 with flags = 0); with the flag a
    constant-coefficient stencil is covered, a matrix with variable coefficients keeps its value
    stream, a coefficient that differs in its last bit (or in its sign) either keeps its group off
    the dictionary (group ids) or becomes one more dictionary entry (one id per thread) -- and y
    is the oracle's bit for bit every time."""
    rng = np.random.default_rng(123 + bits)
    for kind, dims, K in ((E.GEN_LAPLACE2D, (128, 96), 5), (E.GEN_STENCIL27, (24, 20, 16), 27),
                          (E.GEN_STENCIL27, (40, 36, 33), 27)):
        G = E.EllMatrix.generate(kind, dims, (26.0, -1.0), 42, bits, flags=E.NO_PATTERN)
        nr = int(np.prod(dims))
        ec, ea = G.download()
        G.free()
        x = rng.standard_normal(nr)
        y0 = rng.standard_normal(nr)
        variable = rng.standard_normal(ea.shape) * (ea != 0.0)
        # two dents in interior rows (no padded slot): one ulp in one coefficient, the sign of another
        interior = np.flatnonzero((ea.reshape(nr, K) != 0.0).all(axis=1))
        ra, rb = int(interior[len(interior) // 2]), int(interior[len(interior) // 3])
        dented = ea.copy()
        dented[ra * K + 1] = np.nextafter(dented[ra * K + 1], 0.0)
        dented[rb * K + K - 1] = -dented[rb * K + K - 1]
        full = {}
        for name, vals in (("constant", ea), ("variable", variable), ("dented", dented)):
            want = y0.copy()
            oracle.ellgemv(nr, want, x, K, ec, vals)
            w2 = np.zeros(nr)
            oracle.ellgemv(nr, w2, x, K, ec, vals)
            for R in (1, 2):
                for lanes in (0, E.NO_PATTERN_LANES):
                    base = E.EllMatrix.upload(nr, nr, K, ec, vals, E.rows_per_thread(R) | lanes)
                    assert base.info().value_pattern_rows == 0
                    index_rows = base.info().pattern_rows
                    base.free()
                    A = E.EllMatrix.upload(nr, nr, K, ec, vals, E.rows_per_thread(R) | lanes | E.VALUE_PATTERN)
                    info = A.info()
                    key = (R, lanes)
                    if name == "variable":
                        assert info.value_pattern_rows == 0 and info.pattern_rows == index_rows, (kind, dims, key)
                    else:
                        assert info.value_pattern_rows in (0, info.pattern_rows), (kind, dims, key, name)
                        if name == "constant":
                            full[key] = info.value_pattern_rows
                            if index_rows > 0:
                                assert info.value_pattern_rows == index_rows, (kind, dims, key)
                        else:
                            # at most the two groups of the dents leave the dictionary
                            assert full[key] - 2 * 32 * R <= info.value_pattern_rows <= full[key], (kind, dims, key)
                            if lanes and full[key] > 0:
                                # group ids: a group holding a dent stays off the dictionary, if it was on it
                                group = 32 * R
                                on = lambda r: (ea.reshape(nr, K)[r // group * group:(r // group + 1) * group] != 0.0).all()
                                lost = sum(group for g in {ra // group, rb // group} if on(g * group))
                                assert info.value_pattern_rows == full[key] - lost, (kind, dims, key, lost)
                    y = y0.copy()
                    A.spmv(y, x, 1, E.ACCUMULATE)
                    assert bits_equal(y, want), (kind, dims, key, name)
                    y = np.zeros(nr)
                    A.spmv(y, x, 1, E.OVERWRITE)
                    assert bits_equal(y, w2), (kind, dims, key, name, "overwrite")
                    c2, a2 = A.download()
                    assert np.array_equal(c2, ec) and bits_equal(a2, vals)
                    A.free()
        # FMA never searches for value patterns
        A = E.EllMatrix.upload(nr, nr, K, ec, ea, E.VALUE_PATTERN | E.FMA)
        assert A.info().value_pattern_rows == 0
        A.free()




def test_offset_patterns_in_a_row_shard(lib, oracle):
    """A shard's rows are local, its columns global: the pattern offsets are relative to the GLOBAL row."""
    rng = np.random.default_rng(5)
    nr = nc = 3000
    ec, ea = banded_ell(nr, nc, (-40, 0, 1, 40), np.int32, rng)
    x = rng.standard_normal(nc)
    want = np.zeros(nr)
    oracle.ellgemv(nr, want, x, 4, ec, ea)
    lo, hi = 1111, 2777
    A = E.EllMatrix.upload(hi - lo, nc, 4, ec[lo * 4:hi * 4], ea[lo * 4:hi * 4], global_rows=nr, row_begin=lo)
    assert A.info().pattern_rows >= (hi - lo) - 128        # K = 4: two rows per thread, groups of 64
    y = np.zeros(hi - lo)
    A.spmv(y, x, 1, E.OVERWRITE)
    assert bits_equal(y, want[lo:hi])
    A.free()


def expected_pattern_rows(ec, K, nr, R, row_begin=0, max_patterns=16, max_explicit=0, lanes=None, idx_bytes=4,
                          details=False):
    """numpy restatement of the upload-time pattern search (pattern.cu) for matrices small
    enough that every group is sampled.  A group is the 32*R rows of one warp, lane j owning
    rows j*R..j*R+R-1; a lane has an offset vector when its R rows share one; a group's
    signature is the vector at least 32 - max_explicit of its lanes share (max_explicit = 0, the
    default: all of them; 4 with ELLSPMV_CUDA_PATTERN_MASKS); the 16 most common signatures (ties:
    first seen) form the dictionary; in a group whose signature is in it the lanes with another
    vector keep their explicit indices.
    Lane patterns (default when max_explicit = 0, off with NO_PATTERN_LANES): the dictionary is the
    32 most common LANE vectors (ties: lowest thread), a group is patterned when each of its 32
    lanes has its vector in it; taken instead of the group ids when the index bytes of the
    groups it wins exceed twice the byte per thread the ids cost.
    Dropped below 10 % of the groups.
    Returns the rows that take their indices from the dictionary."""
    if lanes is None:
        lanes = max_explicit == 0
    S = 128 * R
    G = 32 * R
    padded = -(-nr // S) * S
    groups = padded // G
    off = ec.reshape(nr, K).astype(np.int64) - (row_begin + np.arange(nr, dtype=np.int64))[:, None]
    sig = {}
    group_sig, group_match, group_lanes = [], [], []
    for g in range(groups):
        lo, hi = g * G, (g + 1) * G
        if hi > nr:
            group_sig.append(None)
            group_match.append(0)
            lv = []
            for j in range(32):
                if lo + (j + 1) * R <= nr:
                    blk = off[lo + j * R: lo + (j + 1) * R]
                    lv.append(tuple(blk[0]) if (blk == blk[0]).all() else None)
                else:
                    lv.append(None)
            group_lanes.append(lv)
            continue
        lv = []
        for j in range(32):
            blk = off[lo + j * R: lo + (j + 1) * R]
            lv.append(tuple(blk[0]) if (blk == blk[0]).all() else None)
        group_lanes.append(lv)
        counts = {}
        for v in lv:
            if v is not None:
                counts[v] = counts.get(v, 0) + 1
        key, cnt = max(counts.items(), key=lambda kv: kv[1]) if counts else (None, 0)
        if cnt >= 32 - max_explicit:
            group_sig.append(key)
            group_match.append(cnt)
            c, first = sig.get(key, (0, g))
            sig[key] = (c + 1, first)
        else:
            group_sig.append(None)
            group_match.append(0)
    best = sorted(sig.items(), key=lambda kv: (-kv[1][0], kv[1][1]))[:max_patterns]
    keep = {k for k, _ in best}
    covered = [g for g in range(groups) if group_sig[g] is not None and group_sig[g] in keep]
    rows = sum(group_match[g] for g in covered) * R
    ncov = len(covered)
    if lanes and max_explicit == 0 and ncov < groups:
        lsig = {}
        for g in range(groups):
            for j, v in enumerate(group_lanes[g]):
                if v is not None:
                    c, first = lsig.get(v, (0, g * 32 + j))
                    lsig[v] = (c + 1, first)
        lkeep = {k for k, _ in sorted(lsig.items(), key=lambda kv: (-kv[1][0], kv[1][1]))[:32]}
        lcov = sum(1 for g in range(groups) if all(v is not None and v in lkeep for v in group_lanes[g]))
        if lkeep and (lcov - ncov) * 32 * R * K * idx_bytes > 2 * groups * 32:
            ncov, rows = lcov, lcov * G
    if ncov * 10 < groups:
        return (0, 0, groups) if details else 0
    return (rows, ncov, groups) if details else rows


@pytest.mark.parametrize("seed", range(8))
def test_offset_patterns_randomized(lib, oracle, seed):
    """Random banded matrices with random damage, every rows-per-thread choice, both index
    widths, whole matrices and row shards: the patterned rows are exactly the ones the numpy
    restatement finds, and y is the oracle's bit for bit."""
    rng = np.random.default_rng(1000 + seed)
    K = int(rng.integers(1, 13))
    nr = int(rng.integers(40, 4000))
    nc = nr + int(rng.integers(0, 50))
    dt = np.int32 if seed % 2 == 0 else np.int64
    offsets = sorted(int(v) for v in rng.integers(-60, 60, K))
    ec, ea = banded_ell(nr, nc, offsets, dt, rng)
    ec = ec.reshape(nr, K)
    for _ in range(int(rng.integers(0, 6))):                       # damage a few entries
        ec[int(rng.integers(0, nr)), int(rng.integers(0, K))] = int(rng.integers(0, nc))
    if seed % 3 == 0:                                               # a second family of rows
        lo = (nr // 2 // 128) * 128
        ec[lo:, 0] = np.minimum(np.arange(lo, nr) + 3, nc - 1)
    ec = ec.reshape(-1)
    x = rng.standard_normal(nc)
    want = np.zeros(nr)
    oracle.ellgemv(nr, want, x, K, ec, ea)
    auto_R = 2 if K <= 12 else 1
    modes = ((0, 0, True), (E.NO_PATTERN_LANES, 0, False), (E.PATTERN_MASKS, 4, False))
    for R in (0, 1, 2, 4):
        for flags, me, lanes in modes:
            A = E.EllMatrix.upload(nr, nc, K, ec, ea, (E.rows_per_thread(R) if R else 0) | flags)
            info = A.info()
            eff_R = R or auto_R
            if R == 0 and auto_R == 2:
                # AUTO's second look (api.cu::build_patterns): two rows per thread only while at least
                # 9 groups in 10 keep their pattern, else the matrix is re-laid with one row per thread
                rows2, cov2, groups2 = expected_pattern_rows(ec, K, nr, 2, max_explicit=me, lanes=lanes, details=True)
                if rows2 > 0 and cov2 * 10 < groups2 * 9:
                    eff_R = 1
            assert info.rows_per_thread == eff_R and info.slice_rows == 128 * eff_R
            assert info.pattern_rows == expected_pattern_rows(ec, K, nr, eff_R, max_explicit=me, lanes=lanes), (seed, K, nr, R, me, lanes)
            y = rng.standard_normal(nr)
            A.spmv(y, x, 1, E.OVERWRITE)
            assert bits_equal(y, want), (seed, K, nr, R, me)
            A.free()
    # a row shard: local rows, global columns
    lo, hi = nr // 5, nr - nr // 7
    for flags, me, lanes in modes:
        A = E.EllMatrix.upload(hi - lo, nc, K, ec[lo * K:hi * K], ea[lo * K:hi * K], E.rows_per_thread(1) | flags,
                               global_rows=nr, row_begin=lo)
        assert A.info().pattern_rows == expected_pattern_rows(ec[lo * K:hi * K], K, hi - lo, 1, row_begin=lo, max_explicit=me,
                                                              lanes=lanes)
        y = np.zeros(hi - lo)
        A.spmv(y, x, 1, E.OVERWRITE)
        assert bits_equal(y, want[lo:hi])
        A.free()
