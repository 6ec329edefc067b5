#!/usr/bin/env python
"""bench.py -- fp64 ELL SpMV throughput on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step is one pass of the hot path (the reference's `ellgemv` loop,
ellspmv.c:1146-1151) over the synthetic 2D 5-point Laplacian on an 8192x8192
grid per GPU (BASELINE config 2: 67,108,864 rows, K = 5, 32-bit indices), in
the y -> x form of BASELINE config 5: x_{k+1} <- A*x_k.  The SAME step is timed
at every N (at N = 1 there is simply nobody to exchange with), so the driver's
scaling efficiency compares like with like.

  value      whole-job GFLOP/s (2*N*K flops per step, padding counted like the
             reference does, ellspmv.c:1857), matrix and vectors resident in HBM
  roofline   the bytes the kernel has to move as the matrix is stored on the
             device (values + the index bytes that are really read + x + y) per
             CUDA-event time, against the measured HBM copy bandwidth
             (MEASURED_PEAKS.json); `algorithmic` is SURVEY.md 8(d)'s figure
             (every index counted at the caller's width), `compression_gain`
             their ratio
  accumulate the reference's own semantics (y <- y + A*x, x constant) on the
             same matrix, kernel only
  e2e        the same metric through the C-ABI call with HOST vectors
             (ellspmv_cuda_spmv, ACCUMULATE: H2D of the x range the shard
             references and of y, launch, D2H of y inside the timing)
  cpu_baseline   the unmodified reference's ellgemv (oracle/_ref) on the box's
             host cores, same matrix -- a reported baseline, not the target
  kernel_switch  (N = 1) the nnz-per-row switch: thread-per-row vs the long-row kernel vs KERNEL_AUTO on
                 2^27-entry random matrices from 4M x 32 to 8192 x 16384, with an on-box bit-parity check
  other_configs  (N = 1) BASELINE configs 3 and 4 (ELL and CSR) and the matrices of
             configs 2 and 3 through the CSR comparison path, kernel only, each
             with its CPU baseline on a bounded sample
  config5    (N > 1) BASELINE config 5 itself: 27-point 768^3, strong-scaled,
             100 iterations, fused push exchange and NCCL all-gather
  parity_check   (N > 1) after k steps every rank recomputes its rows from the
             gathered x_k with the plain single-GPU launch and compares them bit
             for bit with its slice of the sharded x_{k+1}

N > 1 is launched by torchrun, one rank per GPU.  --impl reference times the
reference's own CPU implementation (rank 0 only).
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

# name -> (generator kind, K, index bits, (centre, off) reference values / iterate values, dims(world), scaling)
# iterate values have row sums of at most 1 (|A|_inf <= 1), so x stays O(1) over any number of y -> x steps
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


def workload_name(name: str, world: int) -> str:
    kind, K, bits, _, _, dims_of, _ = WORKLOADS[name]
    return f"{kind}_{'x'.join(str(d) for d in dims_of(world))}_K{K}_idx{bits}"


def algorithmic_bytes(rows: int, ncols: int, K: int, idx_bytes: int, y_rmw: bool) -> int:
    """values + indices + x once + y once (+ y read when y is
    truly read-modify-written, i.e. the reference's accumulate semantics)."""
    return rows * K * (8 + idx_bytes) + 8 * ncols + 8 * rows * (2 if y_rmw else 1)


def kernel_description(info, flags: int) -> str:
    """Which of the library's ELL kernels the handle runs (include/ellspmv_cuda.h)."""
    arith = "fma (tolerance)" if info.fma else "mul-then-add"
    exact = "" if info.fma else " (bit-exact)"
    if getattr(info, "staged", 0):
        return f"staged gather: column blocks, gather staged through HBM, then thread-per-row, {arith}{exact}"
    if flags & (1 << 7):
        return f"column-blocked, per-block partial sums, {arith} (tolerance)"
    if info.kernel == 2:
        return f"sub-warp-per-row + shuffle reduction, {arith} (tolerance)"
    if info.kernel == 4:
        return f"long rows: CTA per row group, products parked in shared memory, sequential adds, {arith}{exact}"
    return f"thread-per-row, {arith}{exact}"


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


def recorded_traffic(key: str):
    """dram__bytes_read+write per launch from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(key)
    except Exception:
        return None


def host_cores() -> dict:
    logical = os.cpu_count() or 1
    try:
        import psutil
        physical = psutil.cpu_count(logical=False) or logical
    except Exception:
        physical = logical
    try:
        usable = len(os.sched_getaffinity(0))
    except Exception:
        usable = logical
    return {"physical": physical, "logical": logical, "usable": usable}


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
            self._stop.clear()
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=1.0)
            self._thread = None

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}

    def nvlink_tx_bytes(self):
        """NVLink payload bytes this GPU has sent so far (NVML throughput counter, KiB), or None."""
        nv = self.nv
        if nv is None:
            return None
        fid = getattr(nv, "NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX", 138)
        for arg in ([(fid, 0xFFFFFFFF)], [fid]):            # all links at once, where the driver takes it
            try:
                v = nv.nvmlDeviceGetFieldValues(self.h, arg)[0]
                if v.nvmlReturn == 0:
                    return int(v.value.ullVal) * 1024
            except Exception:
                pass
        total, ok = 0, False
        for link in range(18):                              # else link by link (NV18 on this box)
            try:
                v = nv.nvmlDeviceGetFieldValues(self.h, [(fid, link)])[0]
                if v.nvmlReturn == 0:
                    total += int(v.value.ullVal) * 1024
                    ok = True
            except Exception:
                break
        return total if ok else None


def bind_near_gpu(torch, index: int) -> dict:
    """Move this rank's host thread onto the CPUs next to its GPU (NVML's ideal affinity) before the
    pinned e2e vectors are allocated and first touched, so that each rank's host buffers live on its
    GPU's NUMA node instead of all ranks sharing one node's memory system."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            p = torch.cuda.get_device_properties(index)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(f"{p.pci_domain_id:08x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0")
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(index)
        before = len(os.sched_getaffinity(0))
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return {"cpus_before": before, "cpus_near_gpu": len(os.sched_getaffinity(0))}
    except Exception as exc:
        return {"error": repr(exc)[:120]}


# ---------------------------------------------------------------------------
# CPU reference arm / cpu_baseline
# ---------------------------------------------------------------------------
def use_all_host_threads() -> int:
    """torchrun injects OMP_NUM_THREADS=1 into every rank; that is not the user's choice of a
    thread count for the CPU baseline.  Give the reference the cores this process may use."""
    n = host_cores()["usable"]
    if os.environ.get("WORLD_SIZE", "1") != "1" and os.environ.get("OMP_NUM_THREADS", "") in ("", "1"):
        os.environ["OMP_NUM_THREADS"] = str(n)
        try:
            import ctypes
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(n)   # in case libgomp is already initialised
        except Exception:
            pass
    return n


def cpu_reference_run(workload: str, world: int, steps: int, warmup: int, budget_s: float, fmt: str = "ell"):
    """Time the reference's own kernel on the host cores, on `workload` as the GPU arm runs it.

    Uses oracle/_ref/libref_{ell,csr}{32,64}.so (the UNMODIFIED reference compiled by
    oracle/Makefile) when present -> kind "reference"; otherwise the oracle's C restatement ->
    kind "port".  The arrays come from the oracle's generator, first-touched in parallel like
    the reference's main() does (Q11); x has the full length of the workload.  A step is a
    bounded sample -- the first `rows` rows -- sized so that (steps + warmup) passes fit the
    budget; GFLOP/s of this bandwidth-bound loop does not depend on the sample length."""
    import numpy as np

    from oracle.pyoracle import Oracle, Reference

    kind_name, K, idx_bits, vals_acc, _, dims_of, _ = WORKLOADS[workload]
    dims = dims_of(world)
    rows_full = dims[0] if kind_name == "random" else int(np.prod(dims))
    ncols = dims[1] if kind_name == "random" else rows_full
    orc = Oracle()
    ref = Reference(fmt, idx_bits) if Reference.available(fmt, idx_bits) else None
    kind = "reference" if ref is not None else "port"
    threads = ref.num_threads() if ref is not None else orc.num_threads()
    idt = np.int32 if idx_bits == 32 else np.int64
    bytes_per_row = K * (8 + idx_bits // 8) + 16

    def build(rows):
        ec = np.empty(rows * K, dtype=idt)
        ea = np.empty(rows * K, dtype=np.float64)
        x = np.empty(ncols, dtype=np.float64)
        y = np.empty(rows, dtype=np.float64)
        if ref is not None and fmt == "ell":
            ref.first_touch(rows, K, ec, ea, ncols, x, y)
        _, _, c, a, _ = orc.gen_ell(kind_name, dims, vals_acc, seed=42, bits=idx_bits, row_begin=0, row_end=rows)
        ec[:] = c
        ea[:] = a
        del c, a
        x[:] = 1.0
        y[:] = 0.0
        if compact:
            # a stencil in CSR form (csr_from_coo keeps the entries of a row in the same order and
            # has no padding): drop the padded slots -- the generated coefficients are non-zero
            keep = ea != 0.0
            state["rowptr"] = np.concatenate(([0], np.cumsum(keep.reshape(rows, K).sum(axis=1, dtype=np.int64))))
            ec, ea = np.ascontiguousarray(ec[keep]), np.ascontiguousarray(ea[keep])
        return ec, ea, x, y

    compact = fmt == "csr" and kind_name != "random"
    state = {"rowptr": None}

    def run(ec, ea, x, y, rows, n):
        if fmt == "csr":
            if compact:
                rowptr = state["rowptr"][: rows + 1]
                lens = np.diff(rowptr)
                kmin, kmax = int(lens.min()), int(lens.max())
            else:
                # every generated row has exactly K stored entries: the CSR arrays are the row-major ELL arrays
                rowptr = np.arange(rows + 1, dtype=np.int64) * K
                kmin = kmax = K
            if ref is not None:
                return ref.csrgemv(rows, y, ncols, x, rowptr, ec, ea, repeat=n, rowsizemin=kmin, rowsizemax=kmax)
            out = []
            for _ in range(n):
                t0 = time.perf_counter()
                orc.csrgemv(rows, y, x, rowptr, ec, ea)
                out.append(time.perf_counter() - t0)
            return np.array(out)
        if ref is not None:
            return ref.ellgemv(rows, y, ncols, x, K, ec, ea, repeat=n)
        out = []
        for _ in range(n):
            t0 = time.perf_counter()
            orc.ellgemv(rows, y, x, K, ec, ea)
            out.append(time.perf_counter() - t0)
        return np.array(out)

    # first guess: 1 GB of matrix, then scale to the budget from a probe
    rows = min(rows_full, max(1 << 16, (1 << 30) // bytes_per_row))
    ec, ea, x, y = build(rows)
    t_probe = float(np.min(run(ec, ea, x, y, rows, 2)))
    want = budget_s / max(steps + warmup + 2, 1)
    if t_probe < want / 2 and rows < rows_full:
        grow = min(rows_full, int(rows * min(want / t_probe, 6.0)))
        if grow > rows * 1.5 and grow * bytes_per_row < (12 << 30):
            rows = grow
            del ec, ea, y
            ec, ea, x, y = build(rows)
    elif t_probe * (steps + warmup) > budget_s and rows > 1 << 18:
        rows = max(1 << 18, int(rows * budget_s / (t_probe * (steps + warmup))))
        nkeep = int(state["rowptr"][rows]) if compact else rows * K
        ec, ea, y = ec[:nkeep], ea[:nkeep], y[:rows]
    run(ec, ea, x, y, rows, max(warmup, 1))
    secs = run(ec, ea, x, y, rows, steps)
    total = float(np.sum(secs))
    stored = int(state["rowptr"][rows]) if compact else rows * K
    flops = 2.0 * stored
    cores = host_cores()
    sample = (f"{steps} timed passes of the reference {'csrgemv' if fmt == 'csr' else 'ellgemv'} over "
              f"{'all' if rows == rows_full else 'the first'} {rows} of {rows_full} rows of {workload_name(workload, world)} "
              f"(x of full length {ncols}), {threads} OpenMP threads on {cores['physical']} physical / "
              f"{cores['logical']} logical cores, after {max(warmup, 1)} warm-up")
    return {
        "kind": kind, "cores": threads, "physical_cores": cores["physical"], "logical_cores": cores["logical"],
        "omp_num_threads": threads, "sample": sample, "rows": rows,
        "gflops": flops * steps / total * 1e-9,
        "best_gflops": flops / float(np.min(secs)) * 1e-9,
        "ms_per_step": total / steps * 1e3,
        "gbs": (stored * (8 + idx_bits // 8) + 16 * rows + (8 * (rows + 1) if fmt == "csr" else 0)
                + (8 * ncols if rows == rows_full else 0)) * steps / total * 1e-9,
    }


def baseline_record(r: dict) -> dict:
    return {"value": round(r["gflops"], 3), "unit": "GFLOP/s", "cores": r["cores"], "kind": r["kind"],
            "physical_cores": r["physical_cores"], "logical_cores": r["logical_cores"],
            "omp_num_threads": r["omp_num_threads"], "sample": r["sample"],
            "best": round(r["best_gflops"], 3), "gbs_effective": round(r["gbs"], 2)}


def workload_config(workload: str, world: int) -> dict:
    """What is computed, not how: identical in the GPU arm and the reference arm."""
    import numpy as np
    kind_name, K, idx_bits, _, _, dims_of, scaling = WORKLOADS[workload]
    dims = dims_of(world)
    rows = dims[0] if kind_name == "random" else int(np.prod(dims))
    return {"workload": workload_name(workload, world), "rows": rows, "rowsize": K, "idx_bits": idx_bits,
            "rows_per_gpu": rows // world, "parallelism": f"rowshard{world}"}


def main_reference(args, rank: int, world: int):
    if rank != 0:
        return 0
    threads = use_all_host_threads()
    r = cpu_reference_run(args.workload, world, args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference",
        "metric": "ell_spmv_fp64_gflops", "value": round(r["gflops"], 3), "unit": "GFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(r["ms_per_step"], 4), "higher_is_better": True,
        "scaling": WORKLOADS[args.workload][6],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "step": "accumulate (y <- y + A*x, x constant): the only form the reference's ellgemv has; timed on a "
                "bounded sample of the rows (see cpu_baseline.sample), host cores only",
        "cpu_baseline": baseline_record(r),
        "e2e": {"value": round(r["gflops"], 3), "unit": "GFLOP/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "host_threads_available": threads,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------
def time_steps(torch, stream, step, steps: int, barrier):
    """CUDA-event time of exactly `steps` calls of step(), barrier + synchronize on both sides."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    barrier()
    return e0.elapsed_time(e1)


def kernel_record(E, info, rows: int, K: int, idx_bits: int, ms: float, y_rmw: bool, x_touched: int, peak: float,
                  entries=None) -> dict:
    """Throughput and the two byte models for one launch that took `ms`.
    as stored: 64-bit indices are kept as 32-bit on the device when they fit, and rows whose
    indices follow an offset pattern never read them."""
    n_entries = rows * K if entries is None else entries
    dev_ib = int(info.dev_idx_bits) // 8 if info is not None else idx_bits // 8
    pattern_rows = int(info.pattern_rows) if info is not None else 0
    alg = n_entries * (8 + idx_bits // 8) + 8 * x_touched + 8 * rows * (2 if y_rmw else 1)
    id_bytes = int(getattr(info, "pattern_id_bytes", 0)) if info is not None else 0
    value_rows = int(getattr(info, "value_pattern_rows", 0)) if info is not None else 0
    stored = (n_entries * (8 + dev_ib) - pattern_rows * K * dev_ib - value_rows * K * 8 + id_bytes
              + 8 * x_touched + 8 * rows * (2 if y_rmw else 1))
    return {"ms_per_step": round(ms, 5), "gflops": round(2.0 * n_entries / ms * 1e-6, 2),
            "as_stored_gbs": round(stored / ms * 1e-6, 1), "frac_as_stored": round(stored / ms * 1e-6 / peak, 4),
            "algorithmic_gbs": round(alg / ms * 1e-6, 1), "frac_algorithmic": round(alg / ms * 1e-6 / peak, 4),
            "bytes_as_stored": stored, "bytes_algorithmic": alg,
            "pattern_rows_frac": round(pattern_rows / max(rows, 1), 4),
            "value_pattern_rows_frac": round(value_rows / max(rows, 1), 4)}


def other_configs(E, torch, dev, sptr, stream, reps: int, peak: float, with_cpu: bool):
    """BASELINE configs 3 and 4 (ELL and the CSR comparison path) and config 2's matrix through the
    CSR path (csrspmv on a structured grid), kernel only, N = 1."""
    import numpy as np
    out = []

    def sync():
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(3):
            fn()
        return time_steps(torch, stream, fn, reps, sync) / reps

    for name, fmt in (("stencil27_384", "ell"), ("random50m", "ell"), ("random50m", "csr"), ("laplace2d", "csr"),
                      ("stencil27_384", "csr")):
        kind_name, K, idx_bits, vals_acc, _, dims_of, _ = WORKLOADS[name]
        kind = {"laplace2d": E.GEN_LAPLACE2D, "stencil27": E.GEN_STENCIL27, "random": E.GEN_RANDOM}[kind_name]
        dims = dims_of(1)
        rows = dims[0] if kind_name == "random" else int(np.prod(dims))
        ncols = dims[1] if kind_name == "random" else rows
        rec = {"config": workload_name(name, 1), "format": fmt, "flags": 0}
        try:
            gen = torch.Generator(device=dev).manual_seed(4321)
            x = torch.randn(ncols, dtype=torch.float64, device=dev, generator=gen)
            y = torch.zeros(rows, dtype=torch.float64, device=dev)
            if fmt == "ell":
                A = E.EllMatrix.generate(kind, dims, vals_acc, 42, idx_bits, flags=0)
                info = A.info()
                ms = timed(lambda: A.spmv_device(y, x, E.ACCUMULATE, sptr))
                rec.update(kernel_record(E, info, rows, K, idx_bits, ms, True, ncols, peak))
                rec["kernel"] = kernel_description(info, 0)
                rec["launches_per_step"] = int(getattr(info, "launches_per_spmv", 1)) or 1
                rec["device_bytes"] = int(info.device_bytes)
            else:
                A = E.CsrMatrix.generate(kind, dims, seed=42, idx_bits=idx_bits, vals=vals_acc)
                ci = A.info()
                nnz = int(ci.csrsize)
                ms = timed(lambda: A.spmv_device(y, x, E.ACCUMULATE, sptr))
                rec.update(kernel_record(E, None, rows, K, idx_bits, ms, True, ncols, peak, entries=nnz))
                # csrspmv.c:2882-2887: the CSR byte model adds the row pointers
                rec["bytes_algorithmic"] += 8 * (rows + 1)
                if ci.ell_view:
                    # what the sliced-ELL view streams: every slot's value (width = the longest row),
                    # the indices of the rows that are not on an offset pattern, the pattern ids, and
                    # one 4-byte length per row when the rows differ (instead of 8-byte row pointers)
                    Kv, dib = int(ci.max_row_len), int(ci.ell_dev_idx_bits) // 8
                    prow = int(ci.ell_pattern_rows)
                    rec["bytes_as_stored"] = (rows * Kv * 8 + (rows - prow) * Kv * dib + int(ci.ell_pattern_id_bytes)
                                              + (4 * rows if ci.ell_view == 1 else 0) + 8 * ncols + 16 * rows)
                    rec["pattern_rows_frac"] = round(prow / max(rows, 1), 4)
                    rec["view"] = {"width": Kv, "slots": rows * Kv, "stored_entries": nnz, "dev_idx_bits": dib * 8,
                                   "row_lengths": ci.ell_view == 1}
                else:
                    rec["bytes_as_stored"] += 8 * (rows + 1)
                rec["as_stored_gbs"] = round(rec["bytes_as_stored"] / ms * 1e-6, 1)
                rec["algorithmic_gbs"] = round(rec["bytes_algorithmic"] / ms * 1e-6, 1)
                rec["frac_as_stored"] = round(rec["as_stored_gbs"] / peak, 4)
                rec["frac_algorithmic"] = round(rec["algorithmic_gbs"] / peak, 4)
                rec["kernel"] = A.describe() if hasattr(A, "describe") else "csr"
                rec["device_bytes"] = int(A.device_bytes())
                if name == "laplace2d":
                    # the CSR host-vector call (csrspmv_cuda_spmv, what INTEGRATION.md's csrspmv.c patch calls)
                    xh = torch.ones(ncols, dtype=torch.float64).pin_memory()
                    yh = torch.zeros(rows, dtype=torch.float64).pin_memory()
                    A.spmv(yh.numpy(), xh.numpy(), 1, E.ACCUMULATE)
                    ts = []
                    for _ in range(3):
                        t0 = time.perf_counter()
                        A.spmv(yh.numpy(), xh.numpy(), 1, E.ACCUMULATE)
                        ts.append(time.perf_counter() - t0)
                    t = sorted(ts)[1]
                    rec["e2e"] = {"ms_per_step": round(t * 1e3, 3), "value": round(2.0 * rows * K / t * 1e-9, 2), "unit": "GFLOP/s",
                                  "h2d_bytes_per_step": (ncols + rows) * 8, "d2h_bytes_per_step": rows * 8,
                                  "api": "csrspmv_cuda_spmv(A, y_host, x_host, 1, ACCUMULATE), pinned host vectors"}
                    del xh, yh
            rec["mode"] = "accumulate (y <- y + A*x)"
            rec["traffic"] = recorded_traffic(f"{rec['config']}_{fmt}")
            A.free()
            del x, y
            torch.cuda.empty_cache()
        except Exception as exc:              # keep the headline line even if a side config fails
            rec["error"] = repr(exc)
        if with_cpu and "error" not in rec:
            try:
                r = cpu_reference_run(name, 1, 3, 1, budget_s=8.0, fmt=fmt)
                rec["cpu_baseline"] = baseline_record(r)
            except Exception as exc:
                rec["cpu_baseline"] = {"value": None, "sample": f"failed: {exc!r}"}
        out.append(rec)
    return out


def pcie_floor(torch, dev, xh, yh, x_touched: int, rows: int, e2e_ms: float) -> dict:
    """The e2e call's own roofline: its H2D bytes (the x range + y) and D2H bytes (y) as plain
    cudaMemcpyAsync between the same pinned host vectors and device buffers, alone and -- like the
    pipelined call -- both directions at once on two streams.  Median of 3, CUDA events."""
    dx = torch.empty(x_touched, dtype=torch.float64, device=dev)
    dy = torch.empty(rows, dtype=torch.float64, device=dev)
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run(do_up: bool, do_down: bool) -> float:
        ts = []
        for _ in range(4):
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record(up)
            down.wait_event(e0)
            if do_up:
                with torch.cuda.stream(up):
                    dx.copy_(xh[:x_touched], non_blocking=True)
                    dy.copy_(yh, non_blocking=True)
            e1.record(up)
            if do_down:
                with torch.cuda.stream(down):
                    yh.copy_(dy, non_blocking=True)
            e2.record(down)
            torch.cuda.synchronize()
            ts.append(max(e0.elapsed_time(e1), e0.elapsed_time(e2)))
        return sorted(ts[1:])[1]

    h2d_bytes, d2h_bytes = (x_touched + rows) * 8, rows * 8
    t_up, t_down, t_both = run(True, False), run(False, True), run(True, True)
    return {"h2d_alone_gbs": round(h2d_bytes / t_up * 1e-6, 1), "d2h_alone_gbs": round(d2h_bytes / t_down * 1e-6, 1),
            "both_directions_ms": round(t_both, 3), "h2d_gbs_while_both": round(h2d_bytes / t_both * 1e-6, 1),
            "floor_ms_per_step": round(t_both, 3), "e2e_over_floor": round(e2e_ms / t_both, 3),
            "how": "plain copies of the step's H2D (x range + y) and D2H (y) bytes between the same pinned vectors and "
                   "device buffers, both directions at once on two streams, no kernel; median of 3"}


def kernel_switch(E, torch, dev, sptr, stream, reps: int, peak: float):
    """The north-star's nnz-per-row switch on the driver's box (N = 1): random ELL matrices of 2^27
    entries from many short rows to few long ones (x small and L2-resident, as in tools/k_sweep.py),
    through KERNEL_AUTO and through each kernel forced; the three results must agree bit for bit."""
    out = []

    def sync():
        torch.cuda.synchronize()

    for rows, K in ((4_194_304, 32), (131_072, 1024), (32_768, 4096), (8_192, 16_384)):
        rec = {"rows": rows, "rowsize": K, "entries": rows * K}
        try:
            ncols = 1 << 20
            gen = torch.Generator(device=dev).manual_seed(99)
            x = torch.randn(ncols, dtype=torch.float64, device=dev, generator=gen)
            ys = {}
            for name, flags in (("auto", 0), ("thread_per_row", E.KERNEL_THREAD | E.rows_per_thread(1)),
                                ("long_row", E.KERNEL_LONGROW)):
                A = E.EllMatrix.generate(E.GEN_RANDOM, (rows, ncols, K), (0.0, 0.0), 42, 32, flags=flags)
                info = A.info()
                y = torch.zeros(rows, dtype=torch.float64, device=dev)

                def fn():
                    A.spmv_device(y, x, E.OVERWRITE, sptr)
                for _ in range(3):
                    fn()
                ms = time_steps(torch, stream, fn, reps, sync) / reps
                nbytes = rows * K * (8 + info.dev_idx_bits // 8) + 8 * rows + 8 * ncols
                rec[name] = {"ms": round(ms, 4), "gbs": round(nbytes / ms * 1e-6, 1), "frac": round(nbytes / ms * 1e-6 / peak, 4)}
                if name == "auto":
                    rec["auto_picks"] = "long_row" if info.kernel == E.KERNEL_LONGROW else "thread_per_row"
                ys[name] = y
                A.free()
            rec["bit_equal"] = bool(torch.equal(ys["auto"], ys["thread_per_row"]) and torch.equal(ys["auto"], ys["long_row"]))
            del ys, x, y
            torch.cuda.empty_cache()
        except Exception as exc:
            rec["error"] = repr(exc)
        out.append(rec)
    return out


def parity_check(E, torch, dist, it, A, sptr, dev) -> dict:
    """x_k gathered on every rank -> plain single-GPU launch over this rank's rows -> compare with
    the slice the sharded step produces from its own (pushed / gathered) copy of x_k."""
    xk = it.gather_result()                                   # full x_k on every rank
    rows = it.hi - it.lo
    want = torch.empty(rows, dtype=torch.float64, device=dev)
    A.spmv_device(want, xk, E.OVERWRITE, sptr)
    it.step(sptr)
    torch.cuda.synchronize()
    it.check()
    got = it.local()
    same = bool(torch.equal(got.view(torch.int64), want.view(torch.int64)))
    finite = bool(torch.isfinite(got).all())
    t = torch.tensor([1 if (same and finite) else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return {"bit_equal": bool(t.item() == 1), "after_steps": it.steps_done - 1,
            "how": "every rank: plain spmv_device over its rows from the all-gathered x_k == its slice of the "
                   "sharded x_{k+1}, fp64 bit patterns, MIN over ranks"}


def run_config5(E, torch, dist, rank, world, local_rank, dev, sptr, stream, peak, sampler, iters: int = 100) -> dict:
    """BASELINE config 5: 27-point stencil 768^3 (IDXTYPEWIDTH=64), rows sharded over the ranks,
    `iters` iterations of x <- A*x, with the fused push exchange and with the NCCL all-gather."""
    import numpy as np

    from ellspmv_b200.sharded import ShardedIterate, partition_rows
    kind_name, K, idx_bits, _, vals_it, dims_of, _ = WORKLOADS["stencil27_768"]
    dims = dims_of(world)
    global_rows = int(np.prod(dims))
    lo, hi = partition_rows(global_rows, world)[rank]
    rows = hi - lo
    rec = {"config": workload_name("stencil27_768", world), "iterations": iters, "scaling": "strong",
           "rows_per_gpu": rows}

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    A = E.EllMatrix.generate(E.GEN_STENCIL27, dims, vals_it, 42, idx_bits, row_begin=lo, row_end=hi, device=local_rank)
    info = A.info()
    x_touched = int(info.max_col - info.min_col + 1)
    rec["kernel"] = kernel_description(info, 0)
    for exchange in ("push", "allgather"):
        it = ShardedIterate(A, rank, world, exchange=exchange)
        it.set_x(lambda a, b: torch.ones(b - a, dtype=torch.float64, device=dev))
        for _ in range(3):
            it.step(sptr)
        barrier()
        tx0 = sampler.nvlink_tx_bytes()
        ms = time_steps(torch, stream, lambda: it.step(sptr), iters, barrier)
        tx1 = sampler.nvlink_tx_bytes()
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item()) / iters
        r = kernel_record(E, info, rows, K, idx_bits, ms, False, x_touched, peak)
        d = it.describe()
        out = {"ms_per_step": r["ms_per_step"], "gflops": round(2.0 * global_rows * K / ms * 1e-6, 1),
               "per_gpu_as_stored_gbs": r["as_stored_gbs"], "per_gpu_frac_as_stored": r["frac_as_stored"],
               "per_gpu_algorithmic_gbs": r["algorithmic_gbs"],
               "bytes_sent_per_step_rank0_planned": d["bytes_sent_per_step_rank0"],
               "bytes_sent_per_step_rank0_nvml": (round((tx1 - tx0) / iters) if tx0 is not None and tx1 is not None else None),
               "barrier": d["barrier"]}
        if exchange == "push":
            out["parity_check"] = parity_check(E, torch, dist, it, A, sptr, dev)
        rec[exchange] = out
        it.close()
    rec["pattern_rows_frac"] = round(int(info.pattern_rows) / max(rows, 1), 4)
    A.free()
    torch.cuda.empty_cache()
    return rec


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

    from ellspmv_b200.sharded import ShardedIterate, partition_rows
    kind_name, K, idx_bits, vals_acc, vals_it, dims_of, scaling = WORKLOADS[args.workload]
    kind = {"laplace2d": E.GEN_LAPLACE2D, "stencil27": E.GEN_STENCIL27, "random": E.GEN_RANDOM}[kind_name]
    dims = dims_of(world)
    global_rows = dims[0] if kind_name == "random" else int(np.prod(dims))
    row_lo, row_hi = partition_rows(global_rows, world)[rank]
    rows = row_hi - row_lo                   # this GPU's rows
    flags = args.flags
    # one matrix for every mode: the iterate values (row sums <= 1) keep x bounded over any number of steps
    A = E.EllMatrix.generate(kind, dims, vals_it if kind_name != "random" else vals_acc, 42, idx_bits,
                             row_begin=row_lo, row_end=row_hi, device=local_rank, flags=flags)
    info = A.info()
    stream = torch.cuda.current_stream()
    sptr = stream.cuda_stream
    peak, peak_src = measured_peak()
    warmup = max(args.warmup, 3)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    x_touched = int(info.max_col - info.min_col + 1) if kind_name != "random" else int(info.num_columns)
    square = int(info.num_columns) == global_rows

    # ---- headline: x_{k+1} <- A*x_k, the same step at every N ----------------------------------
    sharded = None
    if world == 1:
        gen = torch.Generator(device=dev).manual_seed(1234)
        xs = [torch.randn(int(info.num_columns), dtype=torch.float64, device=dev, generator=gen),
              torch.zeros(max(int(info.num_columns), rows), dtype=torch.float64, device=dev)]
        state = {"cur": 0}

        def step():
            c = state["cur"]
            A.spmv_device(xs[1 - c], xs[c], E.OVERWRITE, sptr)
            if square:
                state["cur"] = 1 - c
        step_name = "iterate (x_{k+1} <- A*x_k on two resident vectors; one GPU: no exchange)"
    else:
        sharded = ShardedIterate(A, rank, world, exchange=args.exchange, barrier=args.barrier)
        sharded.set_x(lambda lo, hi: torch.ones(hi - lo, dtype=torch.float64, device=dev))
        step_name = (f"iterate (x_{{k+1}} <- A*x_k, rows sharded over {world} GPUs, exchange={sharded.exchange}, "
                     f"step hand-shake={sharded.barrier if sharded.exchange == 'push' else 'nccl collective'})")

        def step():
            sharded.step(sptr)

    for _ in range(warmup):
        step()
    barrier()
    launches_before = A.info().launches
    sampler.start()
    total_ms = time_steps(torch, stream, step, args.steps, barrier)
    sampler.stop()
    if dist is not None:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    # our kernels launched inside the timed region: the SpMV launches counted by the library, plus one
    # one-warp hand-shake kernel per step (none with --flags FUSED_SYNC, where the SpMV kernel does it)
    timed_launches = int(A.info().launches - launches_before)
    if sharded is not None and sharded.exchange == "push" and sharded.barrier in ("device", "neighbours") \
            and not (flags & E.FUSED_SYNC):
        timed_launches += args.steps

    flops_step = 2.0 * global_rows * K
    value = flops_step / (ms_per_step * 1e-3) * 1e-9
    head = kernel_record(E, info, rows, K, idx_bits, ms_per_step, False, x_touched, peak)
    if sharded is not None:
        sharded.check()

    line = {
        "metric": "ell_spmv_fp64_gflops", "value": round(value, 2), "unit": "GFLOP/s",
        "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": round(ms_per_step, 5), "higher_is_better": True, "scaling": scaling,
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, world),
        "step": step_name,
        "kernel": {"name": kernel_description(info, flags), "rows_per_thread": info.rows_per_thread,
                   "slice_rows": info.slice_rows, "dev_idx_bits": int(info.dev_idx_bits),
                   "pattern_rows_frac": head["pattern_rows_frac"],
                   "value_pattern_rows_frac": head.get("value_pattern_rows_frac", 0.0),   # > 0 only with --flags 0x1000000 (opt-in)
                   "l2": f"inputs larger than L2 ({rows * K * (8 + idx_bits // 8) / 1e9:.1f} GB matrix per GPU vs 126 MB L2), no flush needed"},
        "gbs": round(head["as_stored_gbs"] * world, 1),
        "roofline": {"bound": "hbm", "achieved": head["as_stored_gbs"], "peak": peak, "unit": "GB/s",
                     "frac": head["frac_as_stored"],
                     "traffic": (recorded_traffic(workload_name(args.workload, 1) + "_iterate")
                                 if world == 1 and args.flags == 0 else None),   # the capture is of the default kernel
                     "peak_source": peak_src, "bytes_per_launch": head["bytes_as_stored"],
                     "bytes_model": "as stored on the device: 8*K*rows (values) + index bytes actually read "
                                    "(device index width; rows on an offset pattern read none) + the pattern ids (one byte per group or per thread) + 8*x_touched + 8*rows (y written once)",
                     "algorithmic": {"achieved": head["algorithmic_gbs"], "frac": head["frac_algorithmic"],
                                     "bytes_per_launch": head["bytes_algorithmic"],
                                     "bytes_model": f"SURVEY 8(d): K*(8+{idx_bits // 8})*rows + 8*x_touched + 8*rows"},
                     "compression_gain": round(head["bytes_algorithmic"] / head["bytes_as_stored"], 4)},
        "gpu_launches": timed_launches,
    }

    # ---- the reference's own semantics on the same matrix: y <- y + A*x, x constant ---------------
    if world == 1:
        y = torch.zeros(rows, dtype=torch.float64, device=dev)
        xa = xs[0]
        for _ in range(3):
            A.spmv_device(y, xa, E.ACCUMULATE, sptr)
        n2 = min(args.steps, 50)
        ms2 = time_steps(torch, stream, lambda: A.spmv_device(y, xa, E.ACCUMULATE, sptr), n2, barrier) / n2
        acc = kernel_record(E, info, rows, K, idx_bits, ms2, True, x_touched, peak)
        acc["traffic"] = recorded_traffic(workload_name(args.workload, 1) + "_accumulate") if args.flags == 0 else None
        line["accumulate"] = acc
        del y
    else:
        line["parity_check"] = parity_check(E, torch, dist, sharded, A, sptr, dev)
        line["exchange"] = sharded.describe()
    line["clocks"] = sampler.summary()

    # ---- e2e: the C-ABI call with HOST vectors ------------------------------------------------
    n_e2e = max(1, min(args.steps, args.e2e_steps))
    if sharded is not None:
        sharded.close()
        sharded = None
    if world == 1:
        del xs
    torch.cuda.empty_cache()
    numa = bind_near_gpu(torch, local_rank) if world > 1 else None
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
    # sanity on the host result: A*ones accumulated n times.  With the iterate values (centre 0.5,
    # neighbours 0.125) a row sums to 0.5 + 0.125 * (number of neighbours inside the grid).
    if args.workload == "laplace2d":
        g = yn.reshape(-1, GRID)
        first = rank == 0
        inner = g[1:-1, 1:-1] if world == 1 else g[1 if first else 0:-1 if rank == world - 1 else None, 1:-1]
        assert inner.min() == inner.max() == n_e2e * 1.0, "e2e result is wrong (interior rows)"
        if first:
            assert g[0, 0] == n_e2e * 0.75 and g[0, 1] == n_e2e * 0.875, "e2e result is wrong (corner/edge rows)"
    line["e2e"] = {"value": round(e2e_value, 2), "unit": "GFLOP/s",
                   "h2d_bytes_per_step": (x_touched + rows) * 8, "d2h_bytes_per_step": rows * 8,
                   "steps": n_e2e, "ms_per_step": round(t_e2e / n_e2e * 1e3, 3),
                   "api": "ellspmv_cuda_spmv(A, y_host, x_host, 1, ACCUMULATE) on every rank's shard, pinned host "
                          "vectors; x is uploaded on the column range the shard references only"}
    if numa is not None:
        line["e2e"]["host_thread_affinity"] = numa
    # what bounds the call: the same bytes as plain copies between the same pinned vectors and the
    # device (no kernel), upload and download at the same time on two streams -- the PCIe floor of a step
    try:
        line["e2e"]["pcie"] = pcie_floor(torch, dev, xh, yh, x_touched, rows, t_e2e / n_e2e * 1e3)
    except Exception as exc:
        line["e2e"]["pcie"] = {"error": repr(exc)}
    del xh, yh, xn, yn
    A.free()
    torch.cuda.empty_cache()

    # ---- the other BASELINE configs ---------------------------------------------------------------
    if world == 1 and args.workload == "laplace2d" and not args.no_other_configs:
        line["other_configs"] = other_configs(E, torch, dev, sptr, stream, min(args.steps, 20), peak,
                                              with_cpu=not args.no_cpu_baseline)
        line["kernel_switch"] = kernel_switch(E, torch, dev, sptr, stream, min(args.steps, 10), peak)
    if world > 1 and args.workload == "laplace2d" and not args.no_config5:
        try:
            line["config5"] = run_config5(E, torch, dist, rank, world, local_rank, dev, sptr, stream, peak, sampler,
                                          iters=args.config5_iters)
        except Exception as exc:
            line["config5"] = {"error": repr(exc)}

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            r = cpu_reference_run(args.workload, 1, 5, 1, budget_s=20.0)
            line["cpu_baseline"] = baseline_record(r)
        except Exception as exc:   # the baseline is informative; never lose the GPU line over it
            line["cpu_baseline"] = {"value": None, "unit": "GFLOP/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"failed: {exc!r}"}
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
    ap.add_argument("--barrier", choices=["neighbours", "device", "nccl"], default="neighbours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="laplace2d")
    ap.add_argument("--flags", type=lambda s: int(s, 0), default=0, help="ELLSPMV_CUDA_* upload flags")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--config5-iters", type=int, default=100)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return main_reference(args, rank, world)
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE is 1)")
    return main_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    sys.exit(main())
