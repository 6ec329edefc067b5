#!/usr/bin/env python
"""CSR kernels on a matrix with power-law row lengths (most rows short, a tail of very long
ones): KERNEL_AUTO (SELL-128-sigma + CTA-per-row for the tail) against the native stream,
scalar and vector kernels.  One JSON line per kernel: CUDA-event median per launch, GB/s on the
CSR byte model (csrspmv.c:2882-2887), and whether the result equals the stream kernel's bits.

    python tools/skew_bench.py [--rows 4000000] [--alpha 1.3] [--min-len 4] [--max-len 200000]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402
from bench import measured_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4_000_000)
    ap.add_argument("--alpha", type=float, default=1.3)
    ap.add_argument("--min-len", type=int, default=4)
    ap.add_argument("--max-len", type=int, default=200_000)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    peak, _ = measured_peak()
    rng = np.random.default_rng(42)
    n = args.rows
    lens = np.minimum((args.min_len / rng.random(n) ** (1.0 / args.alpha)).astype(np.int64), args.max_len)
    rowptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(lens, out=rowptr[1:])
    nnz = int(rowptr[-1])
    cols = rng.integers(0, n, nnz, dtype=np.int64).astype(np.int32)
    vals = rng.standard_normal(nnz)
    shape = {"rows": n, "nnz": nnz, "avg_len": round(nnz / n, 2), "max_len": int(lens.max()),
             "rows_over_4096": int((lens > 4096).sum()), "alpha": args.alpha}
    print(json.dumps({"matrix": shape}), flush=True)
    s = torch.cuda.current_stream()
    sptr = s.cuda_stream
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    nbytes = nnz * 12 + 8 * (n + 1) + 8 * n + 16 * n
    ref = None
    for name, flags in (("auto", 0), ("sell", E.KERNEL_CSR_SELL), ("stream", E.KERNEL_THREAD), ("scalar", E.KERNEL_CSR_SCALAR),
                        ("vector (tolerance)", E.KERNEL_WARP)):
        A = E.CsrMatrix.upload(n, n, rowptr, cols, vals, flags)
        y = torch.zeros(n, dtype=torch.float64, device="cuda")
        A.spmv_device(y, x, E.OVERWRITE, sptr)
        torch.cuda.synchronize()
        got = y.clone()
        if name == "stream":
            ref = got
        for _ in range(2):
            A.spmv_device(y, x, E.ACCUMULATE, sptr)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
        ev[0].record(s)
        for i in range(args.reps):
            A.spmv_device(y, x, E.ACCUMULATE, sptr)
            ev[i + 1].record(s)
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))
        ms = ts[len(ts) // 2]
        i = A.info()
        print(json.dumps({"kernel": name, "describe": A.describe(), "ms": round(ms, 4), "gflops": round(2 * nnz / ms * 1e-6, 1),
                          "gbs": round(nbytes / ms * 1e-6, 1), "frac": round(nbytes / ms * 1e-6 / peak, 3),
                          "sell_slots": int(i.sell_slots), "sell_real": int(i.sell_real), "sell_long_rows": int(i.sell_long_rows),
                          "device_MB": round(i.device_bytes / 1e6, 1), "result": got.cpu().numpy().tobytes().__hash__()}), flush=True)
        A.free()
        del y
    print(json.dumps({"note": "equal `result` hashes = identical fp64 bit patterns (auto, sell, stream, scalar must agree)"}), flush=True)


if __name__ == "__main__":
    main()
