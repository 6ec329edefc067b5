#!/usr/bin/env python
"""Kernel-only sweep over BASELINE.json's single-GPU configs and the kernel
variants (rows per thread, thread/sub-warp, FMA, index narrowing, L2 window).
Prints one JSON line per (config, variant): CUDA-event time per launch,
GFLOP/s, effective GB/s (SURVEY.md 8(d) bytes) and fraction of the measured
HBM peak.  Used to pick the defaults recorded in DESIGN.md.

    python tools/bench_configs.py [--configs c2,c3,c4] [--reps 20] [--quick]
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


def time_launches(fn, reps, warmup=3):
    s = torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record(s)
    for i in range(reps):
        fn()
        ev[i + 1].record(s)
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2], ts[0], ev[0].elapsed_time(ev[-1]) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="c2,c3,c4")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--quick", action="store_true", help="small shapes (smoke the script itself)")
    ap.add_argument("--variants", default="all")
    ap.add_argument("--extra-flags", type=lambda t: int(t, 0), default=0, help="OR-ed into every variant's flags")
    ap.add_argument("--no-csr", action="store_true")
    args = ap.parse_args()
    peak, _ = measured_peak()
    q = args.quick
    configs = {
        "c2": ("laplace2d", E.GEN_LAPLACE2D, (1024, 1024) if q else (8192, 8192), (4.0, -1.0), 32),
        "c3": ("stencil27", E.GEN_STENCIL27, (64, 64, 64) if q else (384, 384, 384), (26.0, -1.0), 64),
        "c4": ("random", E.GEN_RANDOM, (500000, 500000, 32) if q else (50_000_000, 50_000_000, 32), (0.0, 0.0), 32),
    }
    sptr = torch.cuda.current_stream().cuda_stream
    for name in args.configs.split(","):
        kind_name, kind, dims, vals, bits = configs[name]
        variants = [("auto", 0), ("auto group-ids (no lane patterns)", E.NO_PATTERN_LANES)]
        for R in (4, 2, 1):
            variants.append((f"thread R={R}", E.KERNEL_THREAD | E.rows_per_thread(R)))
        for R in (1, 2, 4):
            variants.append((f"bulk-async R={R}", E.KERNEL_THREAD | E.rows_per_thread(R) | E.variant(1)))
        if name == "c4":
            variants.append(("staged-gather", E.KERNEL_THREAD | E.STAGED_GATHER))
            variants.append(("column-blocked", E.KERNEL_THREAD | E.COLUMN_BLOCKED))
            variants.append(("column-blocked fma", E.KERNEL_THREAD | E.COLUMN_BLOCKED | E.FMA))
        variants.append(("thread R=4 fma", E.KERNEL_THREAD | E.rows_per_thread(4) | E.FMA))
        variants.append(("thread R=4 l2persist", E.KERNEL_THREAD | E.rows_per_thread(4) | E.L2_PERSIST_X))
        if bits == 64:
            variants.append(("thread R=1 wide-index", E.KERNEL_THREAD | E.rows_per_thread(1) | E.WIDE_INDEX))
            variants.append(("thread R=2 wide-index", E.KERNEL_THREAD | E.rows_per_thread(2) | E.WIDE_INDEX))
        for R in (1, 4):
            variants.append((f"warp S={128 * R}", E.KERNEL_WARP | E.rows_per_thread(R)))
        if args.variants != "all":
            variants = [v for v in variants if any(t in v[0] for t in args.variants.split(","))]
        for vname, flags in variants:
            flags |= args.extra_flags
            if args.extra_flags:
                vname += f" +0x{args.extra_flags:x}"
            A = E.EllMatrix.generate(kind, dims, vals, 42, bits, flags=flags)
            i = A.info()
            rows, ncols, K = i.num_rows, i.num_columns, i.rowsize
            gen = torch.Generator(device="cuda").manual_seed(1)
            x = torch.randn(ncols, dtype=torch.float64, device="cuda", generator=gen)
            y = torch.zeros(rows, dtype=torch.float64, device="cuda")
            for mode_name, mode in (("accumulate", E.ACCUMULATE), ("overwrite", E.OVERWRITE)):
                med, best, avg = time_launches(lambda: A.spmv_device(y, x, mode, sptr), args.reps)
                b = rows * K * (8 + bits // 8) + 8 * ncols + 8 * rows * (2 if mode == E.ACCUMULATE else 1)
                # bytes the kernel really streams: narrowed indices, and none at all for patterned rows
                stored = (rows * K * 8 + (rows - i.pattern_rows) * K * (i.dev_idx_bits // 8) + 8 * ncols
                          + 8 * rows * (2 if mode == E.ACCUMULATE else 1))
                print(json.dumps({"config": name, "kind": kind_name, "dims": dims, "idx_bits": bits, "variant": vname,
                                  "mode": mode_name, "ms_median": round(med, 4), "ms_best": round(best, 4),
                                  "ms_avg": round(avg, 4), "gflops": round(2.0 * rows * K / med * 1e-6, 1),
                                  "gbs_effective": round(b / med * 1e-6, 1), "frac_of_measured_peak": round(b / med * 1e-6 / peak, 4),
                                  "gbs_as_stored": round(stored / med * 1e-6, 1),
                                  "rows_per_thread": i.rows_per_thread,
                                  "pattern_rows_frac": round(i.pattern_rows / max(rows, 1), 4)}), flush=True)
            A.free()
            del x, y
        if name == "c4" and not args.no_csr:
            for vname, flags in (("csr stream (bit-exact)", E.KERNEL_THREAD), ("csr scalar (bit-exact)", 3), ("csr vector T=8", E.KERNEL_WARP)):
                Cm = E.CsrMatrix.generate(E.GEN_RANDOM, dims, 42, bits, flags=flags)
                rows, ncols, K = dims
                gen = torch.Generator(device="cuda").manual_seed(1)
                x = torch.randn(ncols, dtype=torch.float64, device="cuda", generator=gen)
                y = torch.zeros(rows, dtype=torch.float64, device="cuda")
                med, best, avg = time_launches(lambda: Cm.spmv_device(y, x, E.ACCUMULATE, sptr), args.reps)
                b = rows * K * (8 + bits // 8) + 8 * ncols + 16 * rows + 8 * (rows + 1)
                print(json.dumps({"config": name, "kind": "random-csr", "dims": dims, "idx_bits": bits, "variant": vname,
                                  "mode": "accumulate", "ms_median": round(med, 4), "ms_best": round(best, 4),
                                  "ms_avg": round(avg, 4), "gflops": round(2.0 * rows * K / med * 1e-6, 1),
                                  "gbs_effective": round(b / med * 1e-6, 1),
                                  "frac_of_measured_peak": round(b / med * 1e-6 / peak, 4)}), flush=True)
                Cm.free()
                del x, y


if __name__ == "__main__":
    main()
