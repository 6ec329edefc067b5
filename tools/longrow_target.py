#!/usr/bin/env python
"""Few-long-rows target for timing / ncu: a random ELL matrix of `rows` x `K` (x small, L2-resident)
through one chosen kernel.  One JSON line: CUDA-event median per launch.

    python tools/longrow_target.py --rows 32768 --K 4096 [--kernel longrow|thread|auto] [--reps 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=32768)
    ap.add_argument("--K", type=int, default=4096)
    ap.add_argument("--cols", type=int, default=1 << 20)
    ap.add_argument("--kernel", default="longrow", choices=["longrow", "thread", "auto"])
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--idx", type=int, default=32)
    ap.add_argument("--banded", action="store_true", help="consecutive columns per row (coalesced gathers): the HBM-bound case")
    args = ap.parse_args()
    flags = {"longrow": E.KERNEL_LONGROW, "thread": E.KERNEL_THREAD | E.rows_per_thread(1), "auto": 0}[args.kernel]
    s = torch.cuda.current_stream()
    if args.banded:
        import numpy as np
        rng = np.random.default_rng(1)
        base = (np.arange(args.rows, dtype=np.int64) * 37) % args.cols
        cols = ((base[:, None] + np.arange(args.K, dtype=np.int64)[None, :]) % args.cols).astype(np.int32 if args.idx == 32 else np.int64)
        vals = rng.standard_normal(args.rows * args.K)
        A = E.EllMatrix.upload(args.rows, args.cols, args.K, cols.reshape(-1), vals, flags | E.NO_PATTERN)
        del cols, vals
    else:
        A = E.EllMatrix.generate(E.GEN_RANDOM, (args.rows, args.cols, args.K), (0.0, 0.0), 42, args.idx, flags=flags)
    x = torch.randn(args.cols, dtype=torch.float64, device="cuda")
    y = torch.zeros(args.rows, dtype=torch.float64, device="cuda")
    for _ in range(2):
        A.spmv_device(y, x, E.OVERWRITE, s.cuda_stream)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.reps + 1)]
    ev[0].record(s)
    for i in range(args.reps):
        A.spmv_device(y, x, E.OVERWRITE, s.cuda_stream)
        ev[i + 1].record(s)
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.reps))
    ms = ts[len(ts) // 2]
    idx_bytes = A.info().dev_idx_bits // 8
    nbytes = args.rows * args.K * (8 + idx_bytes) + 8 * args.rows + 8 * min(args.cols, args.rows * args.K)
    print(json.dumps({"rows": args.rows, "K": args.K, "kernel": args.kernel, "banded": args.banded, "ms": round(ms, 4),
                      "gbs": round(nbytes / ms / 1e6, 1), "variant": os.environ.get("ELLSPMV_CUDA_LONGROW_VARIANT", ""),
                      "rshift": os.environ.get("ELLSPMV_CUDA_LONGROW_RSHIFT", "")}), flush=True)
    A.free()


if __name__ == "__main__":
    main()
