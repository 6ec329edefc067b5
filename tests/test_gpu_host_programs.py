"""The host programs on a GPU: stdout must equal the unmodified reference's
stdout (recorded in tests/golden/*.json under LC_ALL=C), byte for byte."""
import os
import subprocess

import numpy as np
import pytest

import hostlib
from conftest import GOLDEN_CASES, load_golden, unhex

pytestmark = pytest.mark.gpu


def run(prog, args):
    env = dict(os.environ, LC_ALL="C")
    return subprocess.run([os.path.join(hostlib.BIN, prog)] + args, capture_output=True, text=True, env=env)


@pytest.fixture(scope="module", autouse=True)
def built():
    hostlib.build_host()


def oracle_stdout(oracle, g, bits, passes):
    """What the reference would print (x = ones, y0 = 0) had it not crashed."""
    e = g[f"idx{bits}"]
    dt = np.int32 if bits == 32 else np.int64
    y = np.zeros(g["num_rows"])
    for _ in range(passes):
        oracle.ellgemv(g["num_rows"], y, np.ones(g["num_columns"]), e["rowsize"],
                       np.array(e["ellcolidx"], dtype=dt), unhex(e["ella"]))
    return "%%MatrixMarket vector array real general\n" + f"{len(y)}\n" + "".join("%.15g\n" % v for v in y)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_stdout_equals_reference_stdout(tmp_path, oracle, name):
    g = load_golden(name)
    A = str(tmp_path / "A.mtx")
    hostlib.write_mtx(A, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]), comments=())
    table = {"ellspmv": ("ellspmv", [A]), "ellspmv_repeat2_warmup1": ("ellspmv", ["--repeat=2", "--warmup=1", A]),
             "ellspmv64": ("ellspmv64", [A]), "csrspmv": ("csrspmv", [A]), "csrspmv64": ("csrspmv64", [A])}
    x, y = str(tmp_path / "x.mtx"), str(tmp_path / "y.mtx")
    if "ellspmv_xy" in g["program"]:
        hostlib.write_vec(x, unhex(g["x"]))
        hostlib.write_vec(y, unhex(g["y0"]))
        table["ellspmv_xy"] = ("ellspmv", [A, x, y])
        table["csrspmv_xy"] = ("csrspmv", [A, x, y])
    checked = 0
    for key, (prog, args) in table.items():
        want = g["program"][key]
        r = run(prog, args)
        assert r.returncode == 0, r.stderr
        if want["returncode"] != 0:
            # the reference itself crashes on this input (rows > columns overruns its
            # ellad array, ellspmv.c:1447-1467): expect what its kernel would have printed
            assert key.startswith("ellspmv")
            want = {"stdout": oracle_stdout(oracle, g, 64 if key == "ellspmv64" else 32,
                                            3 if key == "ellspmv_repeat2_warmup1" else 1)}
        assert r.stdout == want["stdout"], (name, key)
        checked += 1
    assert checked >= 5


def test_verbose_lines_have_the_reference_format(tmp_path):
    g = load_golden("rand_square")
    A = str(tmp_path / "A.mtx")
    hostlib.write_mtx(A, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]))
    r = run("ellspmv", ["-v", "--warmup=1", "--repeat=2", "-q", A])
    assert r.returncode == 0 and r.stdout == ""
    lines = r.stderr.splitlines()
    import re
    assert re.match(r"mtxfile_read: \d+\.\d{6} seconds \(\d+\.\d MB/s\)$", lines[0])
    assert re.match(r"ell_from_coo: \d+\.\d{6} seconds, 32 rows, \d+ nonzeros, 12 nonzeros per row$", lines[1])
    gem = [l for l in lines if l.startswith("gemv")]
    assert len(gem) == 3 and gem[0].startswith("gemv (warmup): ")
    for l in gem:
        assert re.match(r"gemv( \(warmup\))?: \d+\.\d{6} seconds \(\d+\.\d{3} Gnz/s, \d+\.\d{3} Gflop/s, \d+\.\d to \d+\.\d GB/s\)$", l), l
    r = run("csrspmv", ["-v", "--partition-nonzeros", A])
    assert r.returncode == 0 and "csr_from_coo: " in r.stderr and "ignored" in r.stderr


def test_synthetic_and_iterate_options(oracle):
    r = run("ellspmv", ["--synthetic=laplace2d:50,40"])
    assert r.returncode == 0
    y = np.array([float(t) for t in r.stdout.splitlines()[2:]])
    i, j = np.divmod(np.arange(2000), 40)
    assert np.array_equal(y, (i == 0).astype(float) + (i == 49) + (j == 0) + (j == 39))
    # y := A^3 x on the scaled 27-point stencil against the oracle's iterate
    K, ncols, ec, ea, _ = oracle.gen_ell("stencil27", (6, 7, 5), (0.5, -1.0 / 52), bits=64)
    want = oracle.ell_iterate(ncols, np.ones(ncols), 3, K, ec, ea)
    r = run("ellspmv64", ["--synthetic=stencil27s:6,7,5", "--iterate", "--repeat=3"])
    assert r.returncode == 0
    assert r.stdout.splitlines()[2:] == ["%.15g" % v for v in want]
    r = run("ellspmv", ["--separate-diagonal", "--synthetic=laplace2d:4,4"])
    assert r.returncode == 1 and "not supported" in r.stderr


def test_csrspmv_synthetic_prints_what_ellspmv_prints():
    """csrspmv --synthetic builds the stencil in CSR form on the device; with x = ones both programs
    print the same row sums (the padded ELL slots add exact zeros), and -v reports the row lengths."""
    for spec, prog_e, prog_c, lens in (("laplace2d:50,40", "ellspmv", "csrspmv", "3 to 5"),
                                       ("stencil27:9,8,7", "ellspmv64", "csrspmv64", "8 to 27"),
                                       ("random:300,200,7", "ellspmv", "csrspmv", "7 to 7")):
        re_ = run(prog_e, [f"--synthetic={spec}"])
        rc = run(prog_c, [f"--synthetic={spec}", "-v", "--repeat=2", "--warmup=1"])
        assert re_.returncode == 0 and rc.returncode == 0, rc.stderr
        if not spec.startswith("random"):
            # two accumulating passes + one warm-up: three times the row sums, exact in fp64
            ye = np.array([float(t) for t in re_.stdout.splitlines()[2:]])
            yc = np.array([float(t) for t in rc.stdout.splitlines()[2:]])
            assert np.array_equal(3.0 * ye, yc)
        assert f"{lens} nonzeros per row" in rc.stderr and rc.stderr.count("gemv") == 3
    r = run("csrspmv", ["--synthetic=laplace2d:4,4", "--sort-rows"])
    assert r.returncode == 1


@pytest.mark.parametrize("name", ["rand_square", "long_rows", "rand_wide"])
def test_ell_sort_rows_prints_what_the_reference_csr_program_prints(tmp_path, name):
    """ellspmv --sort-rows (intended semantics) adds the same products in the same
    order as csrspmv --sort-rows, plus trailing zero padding."""
    g = load_golden(name)
    A = str(tmp_path / "A.mtx")
    hostlib.write_mtx(A, g["num_rows"], g["num_columns"], g["rowidx"], g["colidx"], unhex(g["a"]), comments=())
    r = run("ellspmv", ["--sort-rows", A])
    assert r.returncode == 0, r.stderr
    assert r.stdout.replace("\n-0\n", "\n0\n") == g["program"]["csrspmv_sorted"]["stdout"].replace("\n-0\n", "\n0\n")
