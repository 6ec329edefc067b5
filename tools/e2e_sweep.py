#!/usr/bin/env python
"""The host-vector call (ellspmv_cuda_spmv, ACCUMULATE, pinned vectors) on BASELINE config 2 for
several chunk counts of its upload/compute/download pipeline: wall-clock ms per call, median of 7.
    python tools/e2e_sweep.py > profiles/r2_e2e_chunks.jsonl"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402


def main():
    n = 8192
    A = E.EllMatrix.generate(E.GEN_LAPLACE2D, (n, n), (0.5, 0.125), 42, 32)
    rows = n * n
    xh = torch.ones(rows, dtype=torch.float64).pin_memory()
    yh = torch.zeros(rows, dtype=torch.float64).pin_memory()
    xn, yn = xh.numpy(), yh.numpy()
    for chunks in (0, 8, 16, 24, 32, 48, 64, 128, 0):       # 0: the default plan (chunks ramp up and down)
        if chunks:
            os.environ["ELLSPMV_CUDA_HOST_CHUNKS"] = str(chunks)
        else:
            os.environ.pop("ELLSPMV_CUDA_HOST_CHUNKS", None)
        A.spmv(yn, xn, 1, E.ACCUMULATE)
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            A.spmv(yn, xn, 1, E.ACCUMULATE)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        print(json.dumps({"chunks": chunks or "default (1,1,2,4,8x6,4,2,1,1 of 64 units)", "ms_median": round(ts[3], 3), "ms_min": round(ts[0], 3),
                          "gflops": round(2.0 * rows * 5 / ts[3] * 1e-6, 2)}), flush=True)
    os.environ.pop("ELLSPMV_CUDA_HOST_CHUNKS", None)
    for plan in ("1,1,2,4,8,8,8,8,8,8,4,2,1,1", "1,1,2,4,8,16,16,8,4,2,1,1", "1,1,2,4,8,16,32,32,16,8,4,2,1,1",
                 "1,1,2,4,8,16,16,16,16,16,16,8,4,2,1,1", "1,2,4,8,8,8,8,8,8,4,2,2,1", "2,2,4,8,8,8,8,8,8,4,2,2"):
        os.environ["ELLSPMV_CUDA_HOST_PLAN"] = plan
        A.spmv(yn, xn, 1, E.ACCUMULATE)
        ts = []
        for _ in range(7):
            t0 = time.perf_counter()
            A.spmv(yn, xn, 1, E.ACCUMULATE)
            ts.append((time.perf_counter() - t0) * 1e3)
        ts.sort()
        print(json.dumps({"plan": plan, "units": sum(int(v) for v in plan.split(",")), "ms_median": round(ts[3], 3),
                          "ms_min": round(ts[0], 3), "gflops": round(2.0 * rows * 5 / ts[3] * 1e-6, 2)}), flush=True)
    A.free()


if __name__ == "__main__":
    main()
