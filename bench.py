#!/usr/bin/env python
"""bench.py -- fp64 ELL SpMV throughput on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path (the reference's `ellgemv`, y <- y + A*x,
ellspmv.c:1146-1151) over the synthetic 2D 5-point Laplacian on an 8192x8192
grid per GPU (BASELINE config 2: 67,108,864 rows, K = 5, 32-bit indices).

  value     whole-job GFLOP/s (2*N*K flops per step, padding counted like the
            reference does, ellspmv.c:1857), matrix and vectors resident in HBM
  roofline  algorithmic bytes per launch / CUDA-event time of the kernel
            against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  e2e       the same metric through the C-ABI call with HOST vectors
            (ellspmv_cuda_spmv: H2D x and y, launch, D2H y inside the timing)
  cpu_baseline  the unmodified reference's ellgemv (oracle/_ref) on the box's
            host cores, same matrix -- a reported baseline, not the target

N > 1 (launched by torchrun, one rank per GPU): the grid grows to
(N*8192)x8192 (weak scaling), rows are sharded in contiguous blocks, and a
step is x_{k+1} <- A*x_k with the exchange of y into every rank's next x
(BASELINE config 5's y->x loop).  --impl reference times the reference's own
CPU implementation (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GRID = 8192                      # per-GPU grid is GRID x GRID
K_LAPLACE = 5
IDX_BYTES = 4

# name -> (generator kind, K, index bits, (centre, off) for accumulate / iterate, dims(world), scaling)
# iterate values have row sums of 1 (|A|_inf = 1), so x stays O(1) over any number of y -> x steps
WORKLOADS = {
    # BASELINE config 2, the headline: 8192x8192 grid per GPU, grid grows along x with N (weak)
    "laplace2d": ("laplace2d", 5, 32, (4.0, -1.0), (0.5, 0.125), lambda w: (GRID * w, GRID), "weak"),
    # BASELINE config 3: 384^3 per GPU, IDXTYPEWIDTH=64 (weak along x)
    "stencil27_384": ("stencil27", 27, 64, (26.0, -1.0), (0.5, 1.0 / 52), lambda w: (384 * w, 384, 384), "weak"),
    # BASELINE config 5: 768^3 fixed, row-sharded over N >= 2 GPUs (strong), y -> x
    "stencil27_768": ("stencil27", 27, 64, (26.0, -1.0), (0.5, 1.0 / 52), lambda w: (768, 768, 768), "strong"),
    # BASELINE config 4: random 50M x 32 (N = 1)
    "random50m": ("random", 32, 32, (0.0, 0.0), (0.0, 0.0), lambda w: (50_000_000 * w, 50_000_000 * w, 32), "weak"),
}
FALLBACK_HBM_GBS = 6650.0        # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def algorithmic_bytes(rows: int, ncols: int, K: int, idx_bytes: int, y_rmw: bool) -> int:
    """values + indices + x once + y once (+ y read when y is
    truly read-modify-written, i.e. the reference's accumulate semantics)."""
    return rows * K * (8 + idx_bytes) + 8 * ncols + 8 * rows * (2 if y_rmw else 1)


def kernel_description(flags: int, fma: bool) -> str:
    """Which of the library's ELL kernels the upload flags select (include/ellspmv_cuda.h)."""
    arith = "fma (tolerance)" if fma else "mul-then-add"
    if flags & (1 << 17):
        return f"staged gather: column blocks, gather staged through HBM, then thread-per-row, {arith}" + ("" if fma else " (bit-exact)")
    if flags & (1 << 7):
        return f"column-blocked, per-block partial sums, {arith} (tolerance)"
    if (flags & 0xf) == 2:
        return f"sub-warp-per-row + shuffle reduction, {arith} (tolerance)"
    return f"thread-per-row, {arith}" + ("" if fma else " (bit-exact)")


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(workload: str):
    """dram__bytes_read+write per launch from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed
    region runs (in-process, every few ms: the timed region can be shorter
    than one nvidia-smi period)."""

    def __init__(self, index: int):
        self.index, self.samples, self.reasons = index, [], set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    _NAMES = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
              0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
              0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self._NAMES.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.004)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ---------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ---------------------------------------------------------------------------
def cpu_reference_run(steps: int, warmup: int, budget_s: float):
    """Time the reference's own ellgemv on the host cores, on BASELINE config 2.

    Uses oracle/_ref/libref_ell32.so (the UNMODIFIED reference compiled by
    oracle/Makefile) when present -> kind "reference"; otherwise the oracle's
    C restatement -> kind "port".  The matrix is the same 8192x8192-grid
    Laplacian built by the oracle's generator, first-touched in parallel like
    the reference's main() does (Q11).  If steps*t would exceed the budget,
    a step becomes a bounded sample: the first `sample_rows` rows."""
    import numpy as np

    from oracle.pyoracle import Oracle, Reference

    orc = Oracle()
    ref = Reference("ell", 32) if Reference.available("ell", 32) else None
    kind = "reference" if ref is not None else "port"
    cores = ref.num_threads() if ref is not None else orc.num_threads()
    rows_full = GRID * GRID
    K = K_LAPLACE

    def build(rows):
        ec = np.empty(rows * K, dtype=np.int32)
        ea = np.empty(rows * K, dtype=np.float64)
        x = np.empty(rows_full, dtype=np.float64)
        y = np.empty(rows, dtype=np.float64)
        if ref is not None:
            ref.first_touch(rows, K, ec, ea, rows_full, x, y)
        _, _, c, a, _ = orc.gen_ell("laplace2d", (GRID, GRID), (4.0, -1.0), bits=32, row_begin=0, row_end=rows)
        ec[:] = c
        ea[:] = a
        del c, a
        x[:] = 1.0
        y[:] = 0.0
        return ec, ea, x, y

    def run(ec, ea, x, y, rows, n):
        if ref is not None:
            return ref.ellgemv(rows, y, rows_full, x, K, ec, ea, repeat=n)
        out = []
        for _ in range(n):
            t0 = time.perf_counter()
            orc.ellgemv(rows, y, x, K, ec, ea)
            out.append(time.perf_counter() - t0)
        return np.array(out)

    rows = rows_full
    try:
        ec, ea, x, y = build(rows)
    except MemoryError:
        rows = rows_full // 4
        ec, ea, x, y = build(rows)
    t_probe = float(np.min(run(ec, ea, x, y, rows, 2)))
    if t_probe * (steps + warmup) > budget_s and rows > 1 << 20:
        frac = budget_s / (t_probe * (steps + warmup))
        rows = max(1 << 20, int(rows * frac) // GRID * GRID)
        ec, ea, y = ec[: rows * K], ea[: rows * K], y[:rows]
    run(ec, ea, x, y, rows, max(warmup, 1))
    secs = run(ec, ea, x, y, rows, steps)
    total = float(np.sum(secs))
    flops = 2.0 * rows * K
    sample = (f"{steps} timed passes of the reference ellgemv over "
              f"{'all' if rows == rows_full else 'the first'} {rows} rows of the {GRID}x{GRID}-grid "
              f"5-point Laplacian (K=5, idx32), {cores} OpenMP threads, after {max(warmup, 1)} warm-up")
    return {
        "kind": kind, "cores": cores, "sample": sample, "rows": rows,
        "gflops": flops * steps / total * 1e-9,
        "best_gflops": flops / float(np.min(secs)) * 1e-9,
        "ms_per_step": total / steps * 1e3,
        "gbs": algorithmic_bytes(rows, rows_full if rows == rows_full else rows, K, IDX_BYTES, True) * steps / total * 1e-9,
    }


def main_reference(args, rank: int):
    if rank != 0:
        return 0
    r = cpu_reference_run(args.steps, args.warmup, budget_s=150.0)
    line = {
        "impl": "reference",
        "metric": "ell_spmv_fp64_gflops", "value": round(r["gflops"], 3), "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(r["ms_per_step"], 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"laplace2d_{GRID}x{GRID}_K5_idx32", "mode": "accumulate (y += A*x)",
                   "rows": r["rows"], "host": "CPU, OpenMP"},
        "cpu_baseline": {"value": round(r["gflops"], 3), "unit": "GFLOP/s", "cores": r["cores"],
                         "kind": r["kind"], "sample": r["sample"], "gbs_effective": round(r["gbs"], 2)},
        "e2e": {"value": round(r["gflops"], 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def main_ours(args, rank: int, local_rank: int, world: int):
    import numpy as np
    import torch

    import ellspmv_b200 as E

    if not torch.cuda.is_available() or E.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from ellspmv_b200.sharded import partition_rows
    kind_name, K, idx_bits, vals_acc, vals_it, dims_of, scaling = WORKLOADS[args.workload]
    kind = {"laplace2d": E.GEN_LAPLACE2D, "stencil27": E.GEN_STENCIL27, "random": E.GEN_RANDOM}[kind_name]
    dims = dims_of(world)
    global_rows = dims[0] if kind_name == "random" else int(np.prod(dims))
    row_lo, row_hi = partition_rows(global_rows, world)[rank]
    rows = row_hi - row_lo                   # this GPU's rows
    flags = args.flags
    A = E.EllMatrix.generate(kind, dims, vals_acc if world == 1 else vals_it, 42, idx_bits,
                             row_begin=row_lo, row_end=row_hi, device=local_rank, flags=flags)
    info = A.info()
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)

    if world == 1:
        gen = torch.Generator(device=dev).manual_seed(1234)
        x = torch.randn(int(info.num_columns), dtype=torch.float64, device=dev, generator=gen)
        y = torch.zeros(rows, dtype=torch.float64, device=dev)
        mode_name = "accumulate (y += A*x, the reference's ellgemv semantics)"

        def step():
            A.spmv_device(y, x, E.ACCUMULATE, sptr)
        sharded = None
        y_rmw = True
    else:
        from ellspmv_b200.sharded import ShardedIterate
        sharded = ShardedIterate(A, rank, world, exchange=args.exchange, barrier=args.barrier)
        sharded.set_x(lambda lo, hi: torch.ones(hi - lo, dtype=torch.float64, device=dev))
        mode_name = f"iterate (x <- A*x, exchange={sharded.exchange})"

        def step():
            sharded.step(sptr)
        y_rmw = False

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches_before = A.info().launches
    sampler.start()
    ev[0].record(stream)
    for i in range(args.steps):
        step()
        ev[i + 1].record(stream)
    barrier()
    sampler.stop()
    total_ms = ev[0].elapsed_time(ev[-1])
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    # our kernels launched inside the timed region: the SpMV launches counted by the library,
    # plus one device-barrier kernel per step in the fused push mode
    timed_launches = int(A.info().launches - launches_before)
    if sharded is not None and sharded.exchange == "push" and sharded.barrier == "device":
        timed_launches += args.steps

    flops_step = 2.0 * global_rows * K
    value = flops_step / (ms_per_step * 1e-3) * 1e-9
    # x entries this GPU's launch touches: the column range its rows reference
    x_touched = int(info.max_col - info.min_col + 1) if kind_name != "random" else int(info.num_columns)
    bytes_launch = algorithmic_bytes(rows, x_touched, K, idx_bits // 8, y_rmw)
    bytes_min = algorithmic_bytes(rows, x_touched, K, idx_bits // 8, False)
    # what the kernel really streams: 64-bit indices are stored as 32-bit on the device by default
    # ... and rows whose indices follow an offset pattern do not read their indices at all
    pattern_rows = int(info.pattern_rows)
    bytes_stored = (algorithmic_bytes(rows, x_touched, K, int(info.dev_idx_bits) // 8, y_rmw)
                    - pattern_rows * K * (int(info.dev_idx_bits) // 8))
    peak, peak_src = measured_peak()
    achieved = bytes_launch / (ms_per_step * 1e-3) * 1e-9
    workload = f"{kind_name}_{'x'.join(str(d) for d in dims)}_K{K}_idx{idx_bits}"

    # second kernel-only figure at N=1: overwrite mode (y <- A*x), no y read
    extra = {}
    if world == 1:
        for _ in range(3):
            A.spmv_device(y, x, E.OVERWRITE, sptr)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n2 = min(args.steps, 50)
        e0.record(stream)
        for _ in range(n2):
            A.spmv_device(y, x, E.OVERWRITE, sptr)
        e1.record(stream)
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / n2
        extra["overwrite"] = {"ms_per_step": round(ms2, 5), "gflops": round(flops_step / ms2 * 1e-6, 2),
                              "gbs": round(bytes_min / ms2 * 1e-6, 1), "frac": round(bytes_min / ms2 * 1e-6 / peak, 4)}

    # ---- e2e: the C-ABI call with HOST vectors ----------------------------------
    n_e2e = max(1, min(args.steps, args.e2e_steps))
    xh = torch.empty(int(info.num_columns), dtype=torch.float64).pin_memory()
    yh = torch.zeros(rows, dtype=torch.float64).pin_memory()
    xh.fill_(1.0)
    xn, yn = xh.numpy(), yh.numpy()
    A.spmv(yn, xn, 1, E.ACCUMULATE)          # warm-up: allocates the handle's device vectors
    yn[:] = 0.0
    barrier()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        A.spmv(yn, xn, 1, E.ACCUMULATE)
    t_e2e = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e_value = flops_step * n_e2e / t_e2e * 1e-9
    # sanity: A*ones accumulated n_e2e times is n_e2e * (boundary indicator); checked on the host result
    if world == 1 and args.workload == "laplace2d":
        g = yn.reshape(GRID, GRID)
        assert g[1:-1, 1:-1].max() == 0.0 and g[0, 1] == n_e2e and g[0, 0] == 2 * n_e2e, "e2e result is wrong"

    line = {
        "metric": "ell_spmv_fp64_gflops", "value": round(value, 2), "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload, "rows_per_gpu": rows, "rowsize": K, "idx_bits": idx_bits,
                   "mode": mode_name, "rows_per_thread": info.rows_per_thread, "slice_rows": info.slice_rows,
                   "kernel": kernel_description(flags, bool(info.fma))
                   + (f"; {pattern_rows / max(rows, 1):.1%} of the rows take their column indices from an offset pattern"
                      if pattern_rows else ""),
                   "l2": f"inputs larger than L2 ({rows * K * (8 + idx_bits // 8) / 1e9:.1f} GB matrix per GPU vs 126 MB L2), no flush needed",
                   "parallelism": f"rowshard{world}"},
        "gbs": round(achieved * world, 1),
        "roofline": {"bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                     "frac": round(achieved / peak, 4), "traffic": recorded_traffic((workload + ("_staged" if flags & (1 << 17) else "")) if world == 1 else "sharded"),
                     "peak_source": peak_src, "bytes_per_launch": bytes_launch,
                     "bytes_model": f"K*(8+{idx_bits // 8})*rows + 8*x_touched + 8*rows" + (" + 8*rows (y is read-modify-written)" if y_rmw else ""),
                     "achieved_as_stored": round(bytes_stored / (ms_per_step * 1e-3) * 1e-9, 1),
                     "frac_as_stored": round(bytes_stored / (ms_per_step * 1e-3) * 1e-9 / peak, 4),
                     "dev_idx_bits": int(info.dev_idx_bits),
                     "pattern_rows_frac": round(pattern_rows / max(rows, 1), 4),
                     "achieved_y_once": round(bytes_min / (ms_per_step * 1e-3) * 1e-9, 1),
                     "frac_y_once": round(bytes_min / (ms_per_step * 1e-3) * 1e-9 / peak, 4)},
        "e2e": {"value": round(e2e_value, 2), "unit": "GFLOP/s", "h2d_bytes_per_step": (int(info.num_columns) + rows) * 8,
                "d2h_bytes_per_step": rows * 8, "steps": n_e2e, "ms_per_step": round(t_e2e / n_e2e * 1e3, 3),
                "api": "ellspmv_cuda_spmv(A, y_host, x_host, 1, ACCUMULATE), pinned host vectors"},
        "gpu_launches": timed_launches,
        "clocks": sampler.summary(),
    }
    line.update(extra)
    if sharded is not None:
        line["exchange"] = sharded.describe()

    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload == "laplace2d":
        try:
            r = cpu_reference_run(5, 1, budget_s=25.0)
            line["cpu_baseline"] = {"value": round(r["gflops"], 3), "unit": "GFLOP/s", "cores": r["cores"],
                                    "kind": r["kind"], "sample": r["sample"],
                                    "best": round(r["best_gflops"], 3), "gbs_effective": round(r["gbs"], 2)}
        except Exception as exc:   # the baseline is informative; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc!r}"}
    A.free()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--exchange", choices=["auto", "push", "allgather"], default="auto")
    ap.add_argument("--barrier", choices=["device", "nccl"], default="device")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="laplace2d")
    ap.add_argument("--flags", type=lambda s: int(s, 0), default=0, help="ELLSPMV_CUDA_* upload flags")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return main_reference(args, rank)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE is 1)")
    return main_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
