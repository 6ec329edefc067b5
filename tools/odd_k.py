#!/usr/bin/env python
"""Thread-per-row kernel at row lengths that are not a multiple of the load batch (run-time K path):
random ELL matrices with 2^26 stored entries, x of 2^20 entries (L2-resident), CUDA-event medians.
    python tools/odd_k.py > profiles/r2_odd_k.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402


def main():
    s = torch.cuda.current_stream().cuda_stream
    ncols = 1 << 20
    x = torch.randn(ncols, dtype=torch.float64, device="cuda")
    for K in (3, 5, 6, 7, 8, 9, 11, 12, 13, 15, 16, 17, 20, 24, 27, 31, 32, 33):
        rows = (1 << 26) // K
        for R in (1, 2):
            A = E.EllMatrix.generate(E.GEN_RANDOM, (rows, ncols, K), (0.0, 0.0), 42, 32,
                                     flags=E.KERNEL_THREAD | E.rows_per_thread(R) | E.NO_STAGED_GATHER)
            y = torch.zeros(rows, dtype=torch.float64, device="cuda")
            for _ in range(3):
                A.spmv_device(y, x, E.OVERWRITE, s)
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                A.spmv_device(y, x, E.OVERWRITE, s)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            byts = rows * K * 12 + 8 * rows + 8 * ncols
            print(json.dumps({"K": K, "rows": rows, "rows_per_thread": R, "ms": round(ms, 4),
                              "gbs": round(byts / ms * 1e-6, 1), "gflops": round(2.0 * rows * K / ms * 1e-6, 1)}), flush=True)
            A.free()
            del y


if __name__ == "__main__":
    main()
