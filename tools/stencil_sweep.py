#!/usr/bin/env python
"""Rows-per-thread rule (K <= 12 -> 2 rows per thread, else 1) on stencils OTHER than BASELINE's two:
structured-grid matrices with variable coefficients built as COO on the host, converted on the device
(ellspmv_cuda_upload_coo), run through the thread-per-row kernel with 1 / 2 / 4 rows per thread and
with the AUTO choice.  One JSON line per (stencil, rows per thread): CUDA-event median per launch,
pattern coverage, as-stored GB/s.

    python tools/stencil_sweep.py > profiles/r2_stencil_sweep.jsonl
"""
import itertools
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402
from bench import measured_peak  # noqa: E402


def stencil_coo(dims, offsets, seed=0):
    """COO arrays (1-based, row-major file order, ascending columns inside a row) of a stencil on a
    grid; neighbours outside the grid are left out (so boundary rows are shorter, as in BASELINE's
    generators)."""
    dims = tuple(dims)
    n = int(np.prod(dims))
    strides = [int(np.prod(dims[d + 1:])) for d in range(len(dims))]
    offsets = sorted(offsets, key=lambda o: sum(s * v for s, v in zip(strides, o)))
    coords = np.unravel_index(np.arange(n, dtype=np.int64), dims)
    cols = np.empty((n, len(offsets)), dtype=np.int32)
    for k, o in enumerate(offsets):
        ok = np.ones(n, dtype=bool)
        lin = np.arange(n, dtype=np.int64)
        for d, v in enumerate(o):
            if v:
                c = coords[d] + v
                ok &= (c >= 0) & (c < dims[d])
                lin += v * strides[d]
        cols[:, k] = np.where(ok, lin, -1)
    keep = cols >= 0
    rowidx = np.broadcast_to(np.arange(1, n + 1, dtype=np.int32)[:, None], cols.shape)[keep]
    colidx = (cols[keep] + 1).astype(np.int32)
    vals = np.random.default_rng(seed).standard_normal(len(colidx))
    return n, rowidx.copy(), colidx, vals


def star(dim, radius):
    out = [tuple([0] * dim)]
    for d in range(dim):
        for r in range(1, radius + 1):
            for sgn in (-1, 1):
                o = [0] * dim
                o[d] = sgn * r
                out.append(tuple(o))
    return out


def box(dim, max_nonzero):
    return [o for o in itertools.product((-1, 0, 1), repeat=dim) if sum(1 for v in o if v) <= max_nonzero]


SHAPES = [
    ("star3d_7pt", (224, 224, 224), star(3, 1)),
    ("box2d_9pt", (3072, 3072), box(2, 2)),
    ("star2d_r3_13pt", (3072, 3072), star(2, 3)),
    ("star3d_r2_13pt", (192, 192, 192), star(3, 2)),
    ("box3d_19pt", (176, 176, 176), box(3, 2)),
]


def main():
    peak, _ = measured_peak()
    s = torch.cuda.current_stream()
    for name, dims, offsets in SHAPES:
        n, ri, ci, va = stencil_coo(dims, offsets)
        K = len(offsets)
        x = torch.randn(n, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
        for R in (0, 1, 2, 4):
            A = E.EllMatrix.upload_coo(n, n, ri, ci, va, flags=E.rows_per_thread(R) if R else 0)
            i = A.info()
            y = torch.zeros(n, dtype=torch.float64, device="cuda")
            for _ in range(3):
                A.spmv_device(y, x, E.ACCUMULATE, s.cuda_stream)
            torch.cuda.synchronize()
            reps = 20
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
            ev[0].record(s)
            for r in range(reps):
                A.spmv_device(y, x, E.ACCUMULATE, s.cuda_stream)
                ev[r + 1].record(s)
            torch.cuda.synchronize()
            ts = sorted(ev[r].elapsed_time(ev[r + 1]) for r in range(reps))
            ms = ts[len(ts) // 2]
            prow = int(i.pattern_rows)
            stored = n * K * 8 + (n - prow) * K * (i.dev_idx_bits // 8) + int(i.pattern_id_bytes) + 8 * n + 16 * n
            print(json.dumps({"stencil": name, "dims": list(dims), "rows": n, "K": int(i.rowsize), "entries": len(va),
                              "rows_per_thread": "auto" if R == 0 else R, "picked": int(i.rows_per_thread),
                              "pattern_rows_frac": round(prow / n, 4), "ms": round(ms, 4),
                              "as_stored_gbs": round(stored / ms * 1e-6, 1), "frac_of_measured_peak": round(stored / ms * 1e-6 / peak, 4),
                              "gflops": round(2.0 * n * K / ms * 1e-6, 1)}), flush=True)
            A.free()
            del y
        del x, ri, ci, va


if __name__ == "__main__":
    main()
