#!/usr/bin/env python
"""Short single-GPU target for ncu: builds one BASELINE shape on the device and
launches the chosen kernel a few times.

    python tools/profile_target.py --config c2|c3|c4 [--path ell|csr|csrvec] [--launches 4] [--flags 0x..]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402

CONFIGS = {
    "c2": (E.GEN_LAPLACE2D, (8192, 8192), (4.0, -1.0), 32),
    "c3": (E.GEN_STENCIL27, (384, 384, 384), (26.0, -1.0), 64),
    "c4": (E.GEN_RANDOM, (50_000_000, 50_000_000, 32), (0.0, 0.0), 32),
    "c4s": (E.GEN_RANDOM, (5_000_000, 5_000_000, 32), (0.0, 0.0), 32),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS))
    ap.add_argument("--path", default="ell", choices=["ell", "csr", "csrvec"])  # --flags 0x80 = column-blocked ELL
    ap.add_argument("--launches", type=int, default=4)
    ap.add_argument("--flags", type=lambda s: int(s, 0), default=0)
    ap.add_argument("--mode", default="accumulate", choices=["accumulate", "overwrite"])
    args = ap.parse_args()
    kind, dims, vals, bits = CONFIGS[args.config]
    mode = E.ACCUMULATE if args.mode == "accumulate" else E.OVERWRITE
    s = torch.cuda.current_stream().cuda_stream
    if args.path == "ell":
        A = E.EllMatrix.generate(kind, dims, vals, 42, bits, flags=args.flags)
        i = A.info()
        rows, ncols = i.num_rows, i.num_columns
    else:
        A = E.CsrMatrix.generate(kind, dims, 42, bits, flags=(E.KERNEL_WARP if args.path == "csrvec" else 0) | args.flags,
                                 vals=vals)
        i = A.info()
        rows, ncols = i.num_rows, i.num_columns
        print(A.describe())
    x = torch.randn(ncols, dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    y = torch.zeros(rows, dtype=torch.float64, device="cuda")
    for _ in range(args.launches):
        A.spmv_device(y, x, mode, s)
    torch.cuda.synchronize()
    print("ok", args.config, args.path, float(y[:1000].abs().sum()))
    A.free()


if __name__ == "__main__":
    main()
