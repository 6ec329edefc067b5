#!/usr/bin/env python
"""Generate tests/golden/*.json from the UNMODIFIED reference.

Run in the build container only (needs /root/reference and oracle/_ref/,
i.e. `make -C oracle`).  The reference has no golden vectors of its own
(SURVEY.md 4), so these fixtures are outputs of the reference itself:
  - function level, through oracle/_ref/libref_{ell,csr}{32,64}.so (the
    reference's static ell_from_coo / ellgemv / csr_from_coo / csrgemv);
  - program level, stdout of oracle/_ref/{ellspmv,csrspmv}[64] under LC_ALL=C.
Floating-point vectors are stored as C99 hex floats (bit-exact).
The tests that consume the fixtures never read /root/reference.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.pyoracle import REF_DIR, Reference  # noqa: E402


def hexlist(v):
    return [float(t).hex() for t in v]


def write_mtx(path, nrows, ncols, ri, ci, a, field="real"):
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} general\n")
        f.write(f"{nrows} {ncols} {len(ri)}\n")
        for r, c, v in zip(ri, ci, a):
            if field == "pattern":
                f.write(f"{r} {c}\n")
            else:
                f.write(f"{r} {c} {float(v):.17g}\n")


def write_vec(path, v):
    with open(path, "w") as f:
        f.write("%%MatrixMarket vector array real general\n")
        f.write(f"{len(v)}\n")
        for t in v:
            f.write(f"{float(t):.17g}\n")


def run_binary(name, args):
    env = dict(os.environ, LC_ALL="C", OMP_NUM_THREADS="4")
    p = subprocess.run([os.path.join(REF_DIR, name)] + args, capture_output=True, text=True, env=env)
    return {"returncode": p.returncode, "stdout": p.stdout}


def case(name, nrows, ncols, ri, ci, a, x, y0, tmp):
    out = {"name": name, "num_rows": nrows, "num_columns": ncols,
           "rowidx": [int(t) for t in ri], "colidx": [int(t) for t in ci], "a": hexlist(a),
           "x": hexlist(x), "y0": hexlist(y0)}
    for bits in (32, 64):
        dt = np.int32 if bits == 32 else np.int64
        r, c = np.asarray(ri, dtype=dt), np.asarray(ci, dtype=dt)
        av = np.asarray(a, dtype=np.float64)
        ell = Reference("ell", bits)
        K, ellsize, diagsize, ec, ea = ell.ell_from_coo(nrows, ncols, r, c, av)
        y = np.array(y0, dtype=np.float64)
        ell.ellgemv(nrows, y, ncols, np.asarray(x, dtype=np.float64), K, ec, ea, repeat=1)
        y3 = np.array(y0, dtype=np.float64)
        ell.ellgemv(nrows, y3, ncols, np.asarray(x, dtype=np.float64), K, ec, ea, repeat=3)
        csr = Reference("csr", bits)
        rowptr, cc, ca, lo, hi = csr.csr_from_coo(nrows, ncols, r, c, av)
        yc = np.array(y0, dtype=np.float64)
        csr.csrgemv(nrows, yc, ncols, np.asarray(x, dtype=np.float64), rowptr, cc, ca, 1, lo, hi)
        sc, sa = cc.copy(), ca.copy()
        csr.rowsort(nrows, ncols, rowptr, hi, sc, sa)               # --sort-rows (csrspmv.c:1269-1388)
        ys = np.array(y0, dtype=np.float64)
        csr.csrgemv(nrows, ys, ncols, np.asarray(x, dtype=np.float64), rowptr, sc, sa, 1, lo, hi)
        out[f"idx{bits}"] = {
            "csrcolidx_sorted": [int(t) for t in sc], "csra_sorted": hexlist(sa), "y_csr_sorted": hexlist(ys),
            "rowsize": K, "ellsize": ellsize, "diagsize": diagsize,
            "ellcolidx": [int(t) for t in ec], "ella": hexlist(ea),
            "y_ell": hexlist(y), "y_ell_repeat3": hexlist(y3),
            "rowptr": [int(t) for t in rowptr], "csrcolidx": [int(t) for t in cc], "csra": hexlist(ca),
            "rowsizemin": lo, "rowsizemax": hi, "y_csr": hexlist(yc),
        }
    # whole-program goldens (x = ones, y0 = 0: the reference defaults) + x/y from files
    A = os.path.join(tmp, name + ".mtx")
    write_mtx(A, nrows, ncols, ri, ci, a)
    out["program"] = {
        "ellspmv": run_binary("ellspmv", [A]),
        "ellspmv_repeat2_warmup1": run_binary("ellspmv", ["--repeat=2", "--warmup=1", A]),
        "ellspmv64": run_binary("ellspmv64", [A]),
        "csrspmv": run_binary("csrspmv", [A]),
        "csrspmv64": run_binary("csrspmv64", [A]),
        "csrspmv_sorted": run_binary("csrspmv", ["--sort-rows", A]),
    }
    if nrows == ncols:   # x from file is only right for square A in the reference (Q3)
        xf, yf = os.path.join(tmp, name + "_x.mtx"), os.path.join(tmp, name + "_y.mtx")
        write_vec(xf, x)
        write_vec(yf, y0)
        out["program"]["ellspmv_xy"] = run_binary("ellspmv", [A, xf, yf])
        out["program"]["csrspmv_xy"] = run_binary("csrspmv", [A, xf, yf])
    return out


def case_sd(name, n, ri, ci, a, x, y0, tmp):
    """Separate-diagonal goldens on a square matrix: the reference's functions
    with their flags in DECLARED order (ELL; its main() swaps them, Q1) and the
    CSR program, whose --separate-diagonal path is correct."""
    out = {"name": name, "num_rows": n, "num_columns": n, "rowidx": [int(t) for t in ri],
           "colidx": [int(t) for t in ci], "a": hexlist(a), "x": hexlist(x), "y0": hexlist(y0)}
    xv = np.asarray(x, dtype=np.float64)
    for bits in (32, 64):
        dt = np.int32 if bits == 32 else np.int64
        r, c = np.asarray(ri, dtype=dt), np.asarray(ci, dtype=dt)
        av = np.asarray(a, dtype=np.float64)
        ell = Reference("ell", bits)
        K, ellsize, diagsize, ec, ea, ad = ell.ell_from_coo_sd(n, n, r, c, av)
        y = np.array(y0, dtype=np.float64)
        assert ell.ellgemvsd(0, n, y, n, xv, K, ec, ea, ad) == 0
        y16 = None
        if K == 16:
            y16 = np.array(y0, dtype=np.float64)
            assert ell.ellgemvsd(1, n, y16, n, xv, K, ec, ea, ad) == 0
        csr = Reference("csr", bits)
        rowptr, cc, ca, cad, lo, hi, ds = csr.csr_from_coo_sd(n, n, r, c, av)
        yc = np.array(y0, dtype=np.float64)
        csr.csrgemvsd(n, yc, n, xv, rowptr, cc, ca, cad, lo, hi)
        out[f"idx{bits}"] = {
            "rowsize": K, "ellsize": ellsize, "diagsize": diagsize, "ellcolidx": [int(t) for t in ec],
            "ella": hexlist(ea), "ellad": hexlist(ad), "y_ellgemvsd": hexlist(y),
            "y_ellgemv16sd": hexlist(y16) if y16 is not None else None,
            "rowptr": [int(t) for t in rowptr], "csrcolidx": [int(t) for t in cc], "csra": hexlist(ca),
            "csrad": hexlist(cad), "rowsizemin": lo, "rowsizemax": hi, "csr_diagsize": ds, "y_csrgemvsd": hexlist(yc),
        }
    A = os.path.join(tmp, name + ".mtx")
    write_mtx(A, n, n, ri, ci, a)
    xf, yf = os.path.join(tmp, name + "_x.mtx"), os.path.join(tmp, name + "_y.mtx")
    write_vec(xf, x)
    write_vec(yf, y0)
    out["program"] = {"csrspmv_sd": run_binary("csrspmv", ["--separate-diagonal", A]),
                      "csrspmv64_sd_xy": run_binary("csrspmv64", ["--separate-diagonal", "--repeat=2", A, xf, yf])}
    return out


def main():
    cases = []
    with tempfile.TemporaryDirectory() as tmp:
        # 1. the reference's only fixture, test.mtx (entries parsed here, not copied)
        ri, ci, a = [], [], []
        with open("/root/reference/test.mtx") as f:
            lines = [l for l in f if not l.startswith("%")]
        nrows, ncols, nnz = map(int, lines[0].split())
        for l in lines[1:1 + nnz]:
            r, c, v = l.split()
            ri.append(int(r)); ci.append(int(c)); a.append(float(v))
        cases.append(case("test_mtx", nrows, ncols, ri, ci, a, [1.0] * ncols, [0.0] * nrows, tmp))

        # 2. seeded random matrices: non-square both ways (padding rule Q16),
        #    empty rows, duplicate entries (Q17), unsorted file order
        rng = np.random.default_rng(20261018)
        for name, nr, nc, nnz in [("rand_wide", 23, 37, 140), ("rand_tall", 41, 17, 160),
                                  ("rand_square", 32, 32, 200), ("rand_sparse_rows", 50, 50, 60)]:
            ri = rng.integers(1, nr + 1, nnz)
            ci = rng.integers(1, nc + 1, nnz)
            if name == "rand_sparse_rows":
                ri = 1 + 3 * (ri // 3) % nr   # leaves many rows empty
            ri[5], ci[5] = ri[4], ci[4]       # a duplicate (i, j)
            a = rng.uniform(-2.0, 2.0, nnz)
            x = rng.uniform(-1.0, 1.0, nc)
            y0 = rng.uniform(-1.0, 1.0, nr)
            cases.append(case(name, nr, nc, ri.tolist(), ci.tolist(), a.tolist(), x.tolist(), y0.tolist(), tmp))

        # 3. a single full row and an all-empty matrix
        # a matrix with rows longer than 16 and many repeated columns: exercises the merge tie order
        nr, nc, nnz = 12, 9, 700
        ri = rng.integers(1, nr + 1, nnz); ri[:300] = 3
        ci = rng.integers(1, nc + 1, nnz)
        cases.append(case("long_rows", nr, nc, ri.tolist(), ci.tolist(), rng.uniform(-2, 2, nnz).tolist(),
                          rng.uniform(-1, 1, nc).tolist(), rng.uniform(-1, 1, nr).tolist(), tmp))
        cases.append(case("one_row", 1, 9, [1] * 9, list(range(9, 0, -1)), [float(i) for i in range(1, 10)],
                          [0.5] * 9, [1.0], tmp))
        cases.append(case("empty", 5, 4, [], [], [], [1.0] * 4, [2.0] * 5, tmp))

        # 4. separate-diagonal cases (square): generic, duplicate diagonal entries, K == 16
        rng = np.random.default_rng(77)
        n, nnz = 40, 260
        ri = rng.integers(1, n + 1, nnz); ci = rng.integers(1, n + 1, nnz)
        ri[:30] = ci[:30]                       # plenty of diagonal entries, some repeated
        cases.append(case_sd("sd_rand", n, ri.tolist(), ci.tolist(), rng.uniform(-2, 2, nnz).tolist(),
                             rng.uniform(-1, 1, n).tolist(), rng.uniform(-1, 1, n).tolist(), tmp))
        n = 48                                   # every row: diagonal + exactly 16 off-diagonals -> ellgemv16sd
        ri, ci = [], []
        for i in range(1, n + 1):
            offs = [j for j in rng.permutation(n) + 1 if j != i][:16]
            cols = offs[:7] + [i] + offs[7:]
            ri += [i] * len(cols); ci += [int(j) for j in cols]
        cases.append(case_sd("sd_k16", n, ri, ci, rng.uniform(-2, 2, len(ri)).tolist(),
                             rng.uniform(-1, 1, n).tolist(), rng.uniform(-1, 1, n).tolist(), tmp))

        # 5. a structured grid (2D 5-point stencil) with variable coefficients, entries in the order of
        #    whole groups of 32 / 64 rows share one vector of column offsets, so the
        #    device's offset-pattern path (pattern.cu) is pinned to the reference's own output
        rng = np.random.default_rng(99)
        nx, ny = 3, 200      # grid lines of 200 rows: most groups of 32 / 64 rows lie inside one line
        ri, ci = [], []
        for i in range(nx):
            for j in range(ny):
                r = i * ny + j
                for di, dj in ((-1, 0), (0, -1), (0, 0), (0, 1), (1, 0)):
                    if 0 <= i + di < nx and 0 <= j + dj < ny:
                        ri.append(r + 1); ci.append((i + di) * ny + (j + dj) + 1)
        n = nx * ny
        cases.append(case("grid5", n, n, ri, ci, rng.uniform(-2, 2, len(ri)).tolist(),
                          rng.uniform(-1, 1, n).tolist(), rng.uniform(-1, 1, n).tolist(), tmp))

    for c in cases:
        # tmp paths differ run to run; nothing in stdout depends on them
        with open(os.path.join(HERE, c["name"] + ".json"), "w") as f:
            json.dump(c, f, indent=0, separators=(",", ":"))
        print("wrote", c["name"], {b: c[f"idx{b}"]["rowsize"] for b in (32, 64)},
              [v["returncode"] for v in c["program"].values()])


if __name__ == "__main__":
    main()
