#!/usr/bin/env python
"""Prototype: column-blocked CSR for BASELINE config 4 (random 50M x 32,
x = 400 MB >> L2).  The matrix is split into B column blocks whose x range
fits in L2; y += A_b * x is run block after block with the existing CSR
kernels.  Measures whether keeping the gather inside L2 beats the 28 ms
DRAM-line-fill floor (profiles/r1_c4_gather.md).  Tolerance mode: the
per-row summation order changes."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=50_000_000)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--blocks", default="4,6,8,12")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    N, K = args.rows, args.k
    dims = (N, N, K)
    s = torch.cuda.current_stream().cuda_stream
    x = torch.randn(N, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))

    # reference result and time: the plain ELL kernel
    A = E.EllMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32)
    y_ref = torch.zeros(N, dtype=torch.float64, device="cuda")
    A.spmv_device(y_ref, x, E.OVERWRITE, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        A.spmv_device(y_ref, x, E.OVERWRITE, s)
    e1.record(); torch.cuda.synchronize()
    t_ell = e0.elapsed_time(e1) / args.reps
    A.free()
    print(json.dumps({"variant": "ell thread-per-row (bit-exact)", "ms": round(t_ell, 3)}), flush=True)

    # the CSR arrays of the same matrix, on the device
    Cm = E.CsrMatrix.generate(E.GEN_RANDOM, dims, seed=42, idx_bits=32)
    rowptr = torch.empty(N + 1, dtype=torch.int64, device="cuda")
    cols = torch.empty(N * K, dtype=torch.int32, device="cuda")
    vals = torch.empty(N * K, dtype=torch.float64, device="cuda")
    E._check(E.load_library().csrspmv_cuda_download(Cm._h, rowptr.data_ptr(), cols.data_ptr(), vals.data_ptr()), "download")
    Cm.free()
    del rowptr

    absprod = None
    for B in [int(b) for b in args.blocks.split(",")]:
        W = (N + B - 1) // B
        mats = []
        for b in range(B):
            sel = torch.nonzero((cols >= b * W) & (cols < (b + 1) * W)).squeeze(1)        # keeps (row, file) order
            rows_b = torch.div(sel, K, rounding_mode="floor")
            cnt = torch.bincount(rows_b, minlength=N)
            rp = torch.zeros(N + 1, dtype=torch.int64, device="cuda")
            torch.cumsum(cnt, 0, out=rp[1:])
            cb = cols[sel].contiguous()
            vb = vals[sel].contiguous()
            del sel, rows_b, cnt
            for kern, flags in (("stream", 0), ("vector", E.KERNEL_WARP)):
                mats.append((kern, E.CsrMatrix.upload(N, N, rp, cb, vb, flags)))
            del rp, cb, vb
        for kern in ("stream", "vector"):
            ms_ = [m for k, m in mats if k == kern]
            y = torch.zeros(N, dtype=torch.float64, device="cuda")
            for i, m in enumerate(ms_):
                m.spmv_device(y, x, E.OVERWRITE if i == 0 else E.ACCUMULATE, s)
            torch.cuda.synchronize()
            err = float((y - y_ref).abs().max())
            scale = float(y_ref.abs().max())
            e0.record()
            for _ in range(args.reps):
                for i, m in enumerate(ms_):
                    m.spmv_device(y, x, E.OVERWRITE if i == 0 else E.ACCUMULATE, s)
            e1.record(); torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / args.reps
            print(json.dumps({"variant": f"column-blocked csr {kern}", "blocks": B, "x_block_MB": round(W * 8 / 1e6, 1),
                              "ms": round(t, 3), "speedup_vs_ell": round(t_ell / t, 2), "max_abs_err": err,
                              "max_abs_y": scale}), flush=True)
        for _, m in mats:
            m.free()
        del mats
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
