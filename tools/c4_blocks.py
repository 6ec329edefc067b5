#!/usr/bin/env python
"""BASELINE config 4 (random 50M x 32) through the default path (KERNEL_AUTO -> staged gather) at
several column-block sizes (ELLSPMV_CUDA_BLOCK_BYTES): CUDA-event median of 10 SpMVs each.
    python tools/c4_blocks.py [MB ...] > profiles/r2_c4_block_sweep.jsonl"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import ellspmv_b200 as E  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [16, 24, 32, 40, 48]
    dims = (50_000_000, 50_000_000, 32)
    s = torch.cuda.current_stream().cuda_stream
    x = torch.randn(dims[1], dtype=torch.float64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    y = torch.zeros(dims[0], dtype=torch.float64, device="cuda")
    for mb in sizes:
        os.environ["ELLSPMV_CUDA_BLOCK_BYTES"] = str(mb << 20)
        A = E.EllMatrix.generate(E.GEN_RANDOM, dims, (0.0, 0.0), 42, 32, flags=0)
        info = A.info()
        for _ in range(3):
            A.spmv_device(y, x, E.ACCUMULATE, s)
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            A.spmv_device(y, x, E.ACCUMULATE, s)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        print(json.dumps({"block_mb": mb, "staged": int(info.staged), "launches_per_spmv": int(info.launches_per_spmv),
                          "ms_median": round(ts[len(ts) // 2], 3), "ms_best": round(ts[0], 3),
                          "tune_ms": [round(float(info.tune_ms[0]), 3), round(float(info.tune_ms[1]), 3)],
                          "device_gb": round(info.device_bytes / 1e9, 2)}), flush=True)
        A.free()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
