"""bench.py prints ONE JSON line with the keys the driver's contract names;
checked on a tiny grid (the workload constant is patched) so it runs in seconds."""
import argparse
import io
import json
import contextlib

import pytest

import bench

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"]


def run_main(fn, *args):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        fn(*args)
    lines = [l for l in buf.getvalue().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def test_reference_arm_line(monkeypatch):
    monkeypatch.setattr(bench, "GRID", 96)
    monkeypatch.setitem(bench.WORKLOADS, "laplace2d",
                        ("laplace2d", 5, 32, (4.0, -1.0), (0.5, 0.125), lambda w: (96 * w, 96), "weak"))
    args = argparse.Namespace(gpus=1, steps=3, warmup=1, workload="laplace2d")
    line = run_main(bench.main_reference, args, 0, 1)
    for k in REQUIRED:
        assert k in line, k
    assert line["impl"] == "reference" and line["value"] > 0 and line["unit"] == "GFLOP/s"
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["physical_cores"] >= 1 and line["cpu_baseline"]["omp_num_threads"] >= 1
    # the reference arm describes the workload with the same config keys and values as the GPU arm
    assert line["config"] == bench.workload_config("laplace2d", 1)
    assert line["e2e"] == {"value": line["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("laplace2d_96x96")
    # ranks other than 0 stay silent
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        assert bench.main_reference(args, 1, 2) == 0
    assert buf.getvalue() == ""


def test_reference_arm_under_torchrun_uses_the_host_cores(monkeypatch):
    """torchrun injects OMP_NUM_THREADS=1; the CPU baseline takes the cores the process may use,
    and the matrix is the one the GPU arm runs at that N (N x 8192 x 8192 -> here N x 96 x 96)."""
    monkeypatch.setattr(bench, "GRID", 96)
    monkeypatch.setitem(bench.WORKLOADS, "laplace2d",
                        ("laplace2d", 5, 32, (4.0, -1.0), (0.5, 0.125), lambda w: (96 * w, 96), "weak"))
    monkeypatch.setenv("WORLD_SIZE", "4")
    monkeypatch.setenv("OMP_NUM_THREADS", "1")
    args = argparse.Namespace(gpus=4, steps=2, warmup=1, workload="laplace2d")
    line = run_main(bench.main_reference, args, 0, 4)
    import os
    assert os.environ["OMP_NUM_THREADS"] == str(bench.host_cores()["usable"])
    assert line["config"]["workload"].startswith("laplace2d_384x96") and line["config"] == bench.workload_config("laplace2d", 4)
    assert line["host_threads_available"] == bench.host_cores()["usable"]


def test_algorithmic_bytes_formula():
    # config 2 = 5,100,273,664 B (y once), + 8*rows when y is read-modify-written
    rows = 8192 * 8192
    assert bench.algorithmic_bytes(rows, rows, 5, 4, False) == 5_100_273_664
    assert bench.algorithmic_bytes(rows, rows, 5, 4, True) == 5_100_273_664 + 8 * rows
    assert bench.algorithmic_bytes(384 ** 3, 384 ** 3, 27, 8, False) == 25_367_150_592


@pytest.mark.gpu
def test_gpu_arm_line(monkeypatch, lib):
    monkeypatch.setattr(bench, "GRID", 512)
    monkeypatch.setitem(bench.WORKLOADS, "laplace2d",
                        ("laplace2d", 5, 32, (4.0, -1.0), (0.5, 0.125), lambda w: (512 * w, 512), "weak"))
    args = argparse.Namespace(gpus=1, steps=20, warmup=3, impl="ours", exchange="auto", barrier="fused",
                              workload="laplace2d", flags=0, e2e_steps=2, no_cpu_baseline=False,
                              no_other_configs=True, no_config5=True, config5_iters=100)
    line = run_main(bench.main_ours, args, 0, 0, 1)
    for k in REQUIRED + ["roofline", "cpu_baseline", "clocks", "accumulate", "step", "kernel"]:
        assert k in line, k
    assert line["config"] == bench.workload_config("laplace2d", 1)
    assert line["roofline"]["compression_gain"] >= 1.0 and line["roofline"]["algorithmic"]["frac"] >= line["roofline"]["frac"]
    assert line["accumulate"]["ms_per_step"] > 0
    assert line["gpu_launches"] == 20 and line["n_gpus"] == 1 and line["dtype"] == "f64" and line["vs_baseline"] is None
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    assert line["e2e"]["h2d_bytes_per_step"] == 2 * 512 * 512 * 8 and line["e2e"]["d2h_bytes_per_step"] == 512 * 512 * 8
    assert line["cpu_baseline"]["value"] > 0 and line["cpu_baseline"]["kind"] in ("reference", "port")
