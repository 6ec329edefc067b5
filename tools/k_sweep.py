#!/usr/bin/env python
"""nnz-per-row sweep: thread-per-row (sliced layout) against the long-row kernel (row-major
layout, CTA per row group) at fixed total entries, from many short rows to few long ones.
x is small (1 M columns, L2-resident) so that the matrix stream is what is measured.
One JSON line per (rows, K, kernel): CUDA-event median per launch, GB/s of matrix + vectors.

    python tools/k_sweep.py [--entries 134217728] [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402
from bench import measured_peak  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entries", type=int, default=1 << 27)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--cols", type=int, default=1 << 20)
    args = ap.parse_args()
    peak, _ = measured_peak()
    s = torch.cuda.current_stream()
    sptr = s.cuda_stream
    shapes = []
    for K in (8, 32, 64, 128, 256, 1024, 4096, 16384, 65536, 1 << 20):
        shapes.append((max(args.entries // K, 1), K))
    # few rows at moderate length: the matrix is small, the question is only who fills the GPU
    shapes += [(1024, 1024), (4096, 256), (16384, 64), (256, 4096), (32, 32768), (1, 1 << 22)]
    for rows, K in shapes:
        x = torch.randn(args.cols, dtype=torch.float64, device="cuda")
        y = torch.zeros(rows, dtype=torch.float64, device="cuda")
        for name, flags in (("auto", 0), ("thread", E.KERNEL_THREAD | E.rows_per_thread(1)), ("longrow", E.KERNEL_LONGROW)):
            if name == "thread" and rows * K > (1 << 27) * 2:
                continue
            try:
                A = E.EllMatrix.generate(E.GEN_RANDOM, (rows, args.cols, K), (0.0, 0.0), 42, 32, flags=flags)
            except E.EllspmvCudaError as exc:
                print(json.dumps({"rows": rows, "K": K, "kernel": name, "error": str(exc)}), flush=True)
                continue
            info = A.info()
            for _ in range(2):
                A.spmv_device(y, x, E.OVERWRITE, sptr)
            torch.cuda.synchronize()
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
            ev[0].record(s)
            for i in range(args.reps):
                A.spmv_device(y, x, E.OVERWRITE, sptr)
                ev[i + 1].record(s)
            torch.cuda.synchronize()
            ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))
            ms = ts[len(ts) // 2]
            nbytes = rows * K * 12 + 8 * rows + 8 * min(args.cols, rows * K)
            print(json.dumps({"rows": rows, "K": K, "kernel": name, "picked": int(info.kernel), "slice_rows": int(info.slice_rows),
                              "ms": round(ms, 4), "gbs": round(nbytes / ms * 1e-6, 1), "frac": round(nbytes / ms * 1e-6 / peak, 3),
                              "device_MB": round(info.device_bytes / 1e6, 1)}), flush=True)
            A.free()
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
