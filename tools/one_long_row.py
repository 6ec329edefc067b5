#!/usr/bin/env python
"""Floor of any bit-exact kernel on a long row: the reference adds a row's products one after
another (csrspmv.c:1588-1593), so a row of L entries is L DEPENDENT fp64 additions whatever the
kernel does around them.  Times a CSR matrix of a few rows of L entries each (one CTA per row in
sell.cu's long-row kernel) and prints ns per entry of the longest row.

    python tools/one_long_row.py [--rows 4] [--lens 1000,10000,100000,200000,1000000]
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=4)
    ap.add_argument("--lens", default="1000,10000,100000,200000,1000000")
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    rng = np.random.default_rng(3)
    s = torch.cuda.current_stream()
    for L in (int(t) for t in args.lens.split(",")):
        n, ncols = args.rows, max(L, 1 << 16)
        rowptr = np.arange(n + 1, dtype=np.int64) * L
        cols = rng.integers(0, ncols, n * L).astype(np.int32)
        vals = rng.standard_normal(n * L)
        x = torch.randn(ncols, dtype=torch.float64, device="cuda")
        for name, flags in (("sell (CTA per long row)", E.KERNEL_CSR_SELL), ("stream", E.KERNEL_THREAD)):
            A = E.CsrMatrix.upload(n, ncols, rowptr, cols, vals, flags)
            y = torch.zeros(n, dtype=torch.float64, device="cuda")
            for _ in range(3):
                A.spmv_device(y, x, E.ACCUMULATE, s.cuda_stream)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
            ev[0].record(s)
            for i in range(args.reps):
                A.spmv_device(y, x, E.ACCUMULATE, s.cuda_stream)
                ev[i + 1].record(s)
            torch.cuda.synchronize()
            ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))
            ms = ts[len(ts) // 2]
            print(json.dumps({"rows": n, "row_len": L, "kernel": name, "ms": round(ms, 4),
                              "ns_per_entry_of_one_row": round(ms * 1e6 / L, 2)}), flush=True)
            A.free()


if __name__ == "__main__":
    main()
