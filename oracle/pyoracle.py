"""pyoracle -- ctypes access to the CPU oracle (liboracle.so) and to the
unmodified reference compiled into oracle/_ref/ (libref_*.so).

TEST INFRASTRUCTURE ONLY.  May be imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
legs -- never by the product package (ellspmv_b200/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Tuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_DIR = os.path.join(HERE, "_ref")

_I64 = C.c_int64
_P = C.c_void_p


def build(ref: bool = True) -> None:
    """(Re)build liboracle.so and, if /root/reference is present, oracle/_ref/."""
    subprocess.run(["make", "-C", HERE, "all" if ref else "oracle"], check=True,
                   stdout=subprocess.DEVNULL)


def _p(a: Optional[np.ndarray]):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data


def _idt(bits: int):
    return np.int32 if bits == 32 else np.int64


class Oracle:
    """Our plain-C restatement of the reference algorithms."""

    def __init__(self):
        if not os.path.exists(ORACLE_SO):
            build(ref=False)
        self.lib = C.CDLL(ORACLE_SO)
        self.lib.oracle_num_threads.restype = C.c_int
        self.lib.oracle_splitmix64_export.restype = C.c_uint64
        self.lib.oracle_splitmix64_export.argtypes = [C.c_uint64]

    def _f(self, name: str, bits: int, restype=C.c_int):
        fn = getattr(self.lib, f"{name}{bits}")
        fn.restype = restype
        return fn

    def num_threads(self) -> int:
        return self.lib.oracle_num_threads()

    # -- conversion ---------------------------------------------------------
    def ell_from_coo(self, num_rows: int, num_columns: int, rowidx: np.ndarray, colidx: np.ndarray,
                     a: np.ndarray) -> Tuple[int, int, int, np.ndarray, np.ndarray]:
        """-> (rowsize, ellsize, diagsize, ellcolidx, ella); rowidx/colidx are 1-based."""
        bits = rowidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        nnz = len(a)
        rowcount = np.zeros(num_rows + 1, dtype=np.int64)
        ellsize, rowsize, diagsize = _I64(), I(), I()
        f = self._f("oracle_ell_from_coo_size", bits)
        f.argtypes = [I, I, _I64, _P, _P, C.POINTER(_I64), C.POINTER(I), C.POINTER(I)]
        assert f(num_rows, num_columns, nnz, _p(rowidx), _p(rowcount), C.byref(ellsize), C.byref(rowsize),
                 C.byref(diagsize)) == 0
        ec = np.zeros(ellsize.value, dtype=_idt(bits))
        ea = np.zeros(ellsize.value, dtype=np.float64)
        g = self._f("oracle_ell_from_coo", bits)
        g.argtypes = [I, I, _I64, _P, _P, _P, _P, I, _P, _P]
        assert g(num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowcount), rowsize.value,
                 _p(ec), _p(ea)) == 0
        return rowsize.value, ellsize.value, diagsize.value, ec, ea

    def csr_from_coo(self, num_rows: int, num_columns: int, rowidx: np.ndarray, colidx: np.ndarray,
                     a: np.ndarray):
        """-> (rowptr, csrcolidx, csra, rowsizemin, rowsizemax)"""
        bits = rowidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        nnz = len(a)
        rowptr = np.zeros(num_rows + 1, dtype=np.int64)
        csrsize, lo, hi = _I64(), I(), I()
        f = self._f("oracle_csr_from_coo_size", bits)
        f.argtypes = [I, I, _I64, _P, _P, C.POINTER(_I64), C.POINTER(I), C.POINTER(I)]
        assert f(num_rows, num_columns, nnz, _p(rowidx), _p(rowptr), C.byref(csrsize), C.byref(lo),
                 C.byref(hi)) == 0
        cc = np.zeros(csrsize.value, dtype=_idt(bits))
        ca = np.zeros(csrsize.value, dtype=np.float64)
        g = self._f("oracle_csr_from_coo", bits)
        g.argtypes = [I, _I64, _P, _P, _P, _P, _P, _P]
        assert g(num_rows, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr), _p(cc), _p(ca)) == 0
        return rowptr, cc, ca, lo.value, hi.value

    # -- kernels --------------------------------------------------------------
    def ellgemv(self, num_rows: int, y: np.ndarray, x: np.ndarray, rowsize: int, colidx: np.ndarray,
                a: np.ndarray) -> None:
        bits = colidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        f = self._f("oracle_ellgemv", bits)
        f.argtypes = [I, _P, _P, I, _P, _P]
        assert f(num_rows, _p(y), _p(x), rowsize, _p(colidx), _p(a)) == 0

    def ell_iterate(self, num_rows: int, x: np.ndarray, iterations: int, rowsize: int, colidx: np.ndarray,
                    a: np.ndarray) -> np.ndarray:
        bits = colidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        xa = np.array(x, dtype=np.float64, copy=True)
        xb = np.zeros_like(xa)
        f = self._f("oracle_ell_iterate", bits)
        f.argtypes = [I, _P, _P, C.c_int, I, _P, _P]
        assert f(num_rows, _p(xa), _p(xb), iterations, rowsize, _p(colidx), _p(a)) == 0
        return xa

    def csrgemv(self, num_rows: int, y: np.ndarray, x: np.ndarray, rowptr: np.ndarray, colidx: np.ndarray,
                a: np.ndarray) -> None:
        bits = colidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        f = self._f("oracle_csrgemv", bits)
        f.argtypes = [I, _P, _P, _P, _P, _P]
        assert f(num_rows, _p(y), _p(x), _p(rowptr), _p(colidx), _p(a)) == 0

    # -- separate-diagonal variants ------------------------------------------------
    def ell_from_coo_sd(self, num_rows, num_columns, rowidx, colidx, a):
        """-> (rowsize, ellsize, diagsize, ellcolidx, ella, ellad): flags in declared order (Q1)."""
        bits = rowidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        nnz = len(a)
        rowcount = np.zeros(num_rows + 1, dtype=np.int64)
        ellsize, rowsize, diagsize = _I64(), I(), I()
        f = self._f("oracle_ell_from_coo_sd_size", bits)
        f.argtypes = [I, I, _I64, _P, _P, _P, C.POINTER(_I64), C.POINTER(I), C.POINTER(I)]
        assert f(num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(rowcount), C.byref(ellsize),
                 C.byref(rowsize), C.byref(diagsize)) == 0
        ec = np.zeros(ellsize.value, dtype=_idt(bits))
        ea = np.zeros(ellsize.value, dtype=np.float64)
        ad = np.zeros(max(diagsize.value, 1), dtype=np.float64)[:diagsize.value]
        g = self._f("oracle_ell_from_coo_sd", bits)
        g.argtypes = [I, I, _I64, _P, _P, _P, _P, I, _P, _P, _P]
        assert g(num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowcount), rowsize.value,
                 _p(ec), _p(ea), _p(ad)) == 0
        return rowsize.value, ellsize.value, diagsize.value, ec, ea, ad

    def ellgemvsd(self, num_rows, y, x, rowsize, colidx, a, ad, order: int = 0) -> None:
        bits = colidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        f = self._f("oracle_ellgemvsd", bits)
        f.argtypes = [I, _P, _P, I, _P, _P, _P, C.c_int]
        assert f(num_rows, _p(y), _p(x), rowsize, _p(colidx), _p(a), _p(ad), order) == 0

    def csr_from_coo_sd(self, num_rows, rowidx, colidx, a):
        """square general matrix -> (rowptr, csrcolidx, csra, csrad, rowsizemin, rowsizemax)"""
        bits = rowidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        nnz = len(a)
        rowptr = np.zeros(num_rows + 1, dtype=np.int64)
        csrsize, lo, hi = _I64(), I(), I()
        f = self._f("oracle_csr_from_coo_sd", bits)
        f.argtypes = [I, _I64, _P, _P, _P, _P, C.POINTER(_I64), C.POINTER(I), C.POINTER(I), _P, _P, _P]
        assert f(num_rows, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr), C.byref(csrsize), C.byref(lo),
                 C.byref(hi), None, None, None) == 0
        cc = np.zeros(max(csrsize.value, 1), dtype=_idt(bits))[:csrsize.value]
        ca = np.zeros(max(csrsize.value, 1), dtype=np.float64)[:csrsize.value]
        ad = np.zeros(max(num_rows, 1), dtype=np.float64)[:num_rows]
        assert f(num_rows, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr), C.byref(csrsize), C.byref(lo),
                 C.byref(hi), _p(cc), _p(ca), _p(ad)) == 0
        return rowptr, cc, ca, ad, lo.value, hi.value

    def csrgemvsd(self, num_rows, y, x, rowptr, colidx, a, ad) -> None:
        bits = colidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        f = self._f("oracle_csrgemvsd", bits)
        f.argtypes = [I, _P, _P, _P, _P, _P, _P]
        assert f(num_rows, _p(y), _p(x), _p(rowptr), _p(colidx), _p(a), _p(ad)) == 0

    def rowsort(self, num_rows, rowptr, colidx, a) -> None:
        """--sort-rows on a CSR matrix, in place (csrspmv.c:1269-1388)."""
        bits = colidx.dtype.itemsize * 8
        I = C.c_int32 if bits == 32 else C.c_int64
        f = self._f("oracle_rowsort", bits)
        f.argtypes = [I, _P, _P, _P]
        assert f(num_rows, _p(rowptr), _p(colidx), _p(a)) == 0

    # -- synthetic matrices -----------------------------------------------------
    def gen_ell(self, kind: str, dims, vals=(0.0, 0.0), seed: int = 42, bits: int = 32,
                row_begin: int = 0, row_end: Optional[int] = None):
        """Row-major ELL arrays of rows [row_begin, row_end) -> (K, ncols, colidx, a, real_nnz)."""
        if kind == "laplace2d":
            nx, ny = dims[:2]
            rows, K = nx * ny, 5
        elif kind == "stencil27":
            nx, ny, nz = dims[:3]
            rows, K = nx * ny * nz, 27
        elif kind == "random":
            rows, ncols, K = dims[:3]
        else:
            raise ValueError(kind)
        if row_end is None:
            row_end = rows
        n = (row_end - row_begin) * K
        ec = np.empty(n, dtype=_idt(bits))
        ea = np.empty(n, dtype=np.float64)
        if kind == "laplace2d":
            f = self._f("oracle_gen_laplace2d_ell", bits, _I64)
            f.argtypes = [_I64, _I64, C.c_double, C.c_double, _I64, _I64, _P, _P]
            real = f(nx, ny, vals[0], vals[1], row_begin, row_end, _p(ec), _p(ea))
            ncols = rows
        elif kind == "stencil27":
            f = self._f("oracle_gen_stencil27_ell", bits, _I64)
            f.argtypes = [_I64, _I64, _I64, C.c_double, C.c_double, _I64, _I64, _P, _P]
            real = f(nx, ny, nz, vals[0], vals[1], row_begin, row_end, _p(ec), _p(ea))
            ncols = rows
        else:
            f = self._f("oracle_gen_random_ell", bits, _I64)
            f.argtypes = [_I64, _I64, _I64, C.c_uint64, _I64, _I64, _P, _P]
            real = f(rows, ncols, K, seed, row_begin, row_end, _p(ec), _p(ea))
        return K, ncols, ec, ea, real

    def gen_coo(self, kind: str, dims, vals=(0.0, 0.0), seed: int = 42, bits: int = 32):
        """1-based COO stream in canonical order -> (rows, ncols, rowidx, colidx, a)."""
        if kind == "laplace2d":
            nx, ny = dims[:2]
            rows = ncols = nx * ny
            f = self._f("oracle_gen_laplace2d_coo", bits, _I64)
            f.argtypes = [_I64, _I64, C.c_double, C.c_double, _P, _P, _P]
            args = (nx, ny, vals[0], vals[1])
        elif kind == "stencil27":
            nx, ny, nz = dims[:3]
            rows = ncols = nx * ny * nz
            f = self._f("oracle_gen_stencil27_coo", bits, _I64)
            f.argtypes = [_I64, _I64, _I64, C.c_double, C.c_double, _P, _P, _P]
            args = (nx, ny, nz, vals[0], vals[1])
        elif kind == "random":
            rows, ncols, K = dims[:3]
            f = self._f("oracle_gen_random_coo", bits, _I64)
            f.argtypes = [_I64, _I64, _I64, C.c_uint64, _P, _P, _P]
            args = (rows, ncols, K, seed)
        else:
            raise ValueError(kind)
        nnz = f(*args, None, None, None)
        ri = np.empty(nnz, dtype=_idt(bits))
        ci = np.empty(nnz, dtype=_idt(bits))
        a = np.empty(nnz, dtype=np.float64)
        assert f(*args, _p(ri), _p(ci), _p(a)) == nnz
        return rows, ncols, ri, ci, a


class Reference:
    """The UNMODIFIED reference functions, compiled from /root/reference into
    oracle/_ref/ by oracle/Makefile (kind = 'ell' | 'csr', bits = 32 | 64)."""

    def __init__(self, kind: str, bits: int):
        path = os.path.join(REF_DIR, f"libref_{kind}{bits}.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.kind, self.bits = kind, bits
        self.lib = C.CDLL(path)
        assert self.lib.ref_idx_bytes() * 8 == bits
        self.lib.ref_num_threads.restype = C.c_int

    @staticmethod
    def available(kind: str = "ell", bits: int = 32) -> bool:
        return os.path.exists(os.path.join(REF_DIR, f"libref_{kind}{bits}.so"))

    def num_threads(self) -> int:
        return self.lib.ref_num_threads()

    def ell_from_coo(self, num_rows, num_columns, rowidx, colidx, a):
        nnz = len(a)
        rowptr = np.zeros(num_rows + 1, dtype=np.int64)
        ellsize, rowsize, diagsize = _I64(), _I64(), _I64()
        f = self.lib.ref_ell_from_coo_size
        f.argtypes = [_I64, _I64, _I64, _P, _P, _P, _P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]
        assert f(num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr), C.byref(ellsize),
                 C.byref(rowsize), C.byref(diagsize)) == 0
        ec = np.empty(ellsize.value, dtype=_idt(self.bits))
        ea = np.empty(ellsize.value, dtype=np.float64)
        ead = np.empty(max(diagsize.value, 1), dtype=np.float64)
        g = self.lib.ref_ell_from_coo
        g.argtypes = [_I64, _I64, _I64, _P, _P, _P, _P, _I64, _I64, _P, _P, _P]
        assert g(num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr), ellsize.value,
                 rowsize.value, _p(ec), _p(ea), _p(ead)) == 0
        return rowsize.value, ellsize.value, diagsize.value, ec, ea

    def ellgemv(self, num_rows, y, num_columns, x, rowsize, colidx, a, repeat: int = 1) -> np.ndarray:
        secs = np.zeros(repeat, dtype=np.float64)
        f = self.lib.ref_ellgemv
        f.argtypes = [_I64, _P, _I64, _P, _I64, _I64, _P, _P, C.c_int, _P]
        assert f(num_rows, _p(y), num_columns, _p(x), num_rows * rowsize, rowsize, _p(colidx), _p(a), repeat,
                 _p(secs)) == 0
        return secs

    def first_touch(self, num_rows, rowsize, colidx, a, num_columns, x, y) -> None:
        f = self.lib.ref_first_touch
        f.restype = None
        f.argtypes = [_I64, _I64, _P, _P, _I64, _P, _P]
        f(num_rows, rowsize, _p(colidx), _p(a), num_columns, _p(x), _p(y))

    def csr_from_coo(self, num_rows, num_columns, rowidx, colidx, a, symmetric: bool = False):
        nnz = len(a)
        rowptr = np.zeros(num_rows + 1, dtype=np.int64)
        csrsize, lo, hi = _I64(), _I64(), _I64()
        f = self.lib.ref_csr_from_coo_size
        f.argtypes = [C.c_int, _I64, _I64, _I64, _P, _P, _P, _P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]
        assert f(int(symmetric), num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr),
                 C.byref(csrsize), C.byref(lo), C.byref(hi)) == 0
        cc = np.empty(max(csrsize.value, 1), dtype=_idt(self.bits))[:csrsize.value]
        ca = np.empty(max(csrsize.value, 1), dtype=np.float64)[:csrsize.value]
        g = self.lib.ref_csr_from_coo
        g.argtypes = [C.c_int, _I64, _I64, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P]
        assert g(int(symmetric), num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr),
                 csrsize.value, lo.value, hi.value, _p(cc), _p(ca)) == 0
        return rowptr, cc, ca, lo.value, hi.value

    def csrgemv(self, num_rows, y, num_columns, x, rowptr, colidx, a, repeat: int = 1,
                rowsizemin: int = 0, rowsizemax: int = 0) -> np.ndarray:
        secs = np.zeros(repeat, dtype=np.float64)
        f = self.lib.ref_csrgemv
        f.argtypes = [_I64, _P, _I64, _P, _I64, _I64, _I64, _P, _P, _P, C.c_int, _P]
        assert f(num_rows, _p(y), num_columns, _p(x), int(rowptr[num_rows]), rowsizemin, rowsizemax,
                 _p(rowptr), _p(colidx), _p(a), repeat, _p(secs)) == 0
        return secs

    # -- separate-diagonal functions of the reference -------------------------------
    def ell_from_coo_sd(self, num_rows, num_columns, rowidx, colidx, a):
        nnz = len(a)
        rowptr = np.zeros(num_rows + 1, dtype=np.int64)
        ellsize, rowsize, diagsize = _I64(), _I64(), _I64()
        f = self.lib.ref_ell_from_coo_sd_size
        f.argtypes = [_I64, _I64, _I64, _P, _P, _P, _P, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(_I64)]
        assert f(num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr), C.byref(ellsize),
                 C.byref(rowsize), C.byref(diagsize)) == 0
        ec = np.empty(max(ellsize.value, 1), dtype=_idt(self.bits))[:ellsize.value]
        ea = np.empty(max(ellsize.value, 1), dtype=np.float64)[:ellsize.value]
        ad = np.empty(max(diagsize.value, 1), dtype=np.float64)[:diagsize.value]
        g = self.lib.ref_ell_from_coo_sd
        g.argtypes = [_I64, _I64, _I64, _P, _P, _P, _P, _I64, _I64, _P, _P, _P]
        assert g(num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr), ellsize.value,
                 rowsize.value, _p(ec), _p(ea), _p(ad)) == 0
        return rowsize.value, ellsize.value, diagsize.value, ec, ea, ad

    def ellgemvsd(self, which, num_rows, y, num_columns, x, rowsize, colidx, a, ad) -> int:
        f = self.lib.ref_ellgemvsd
        f.argtypes = [C.c_int, _I64, _P, _I64, _P, _I64, _I64, _P, _P, _P]
        return f(which, num_rows, _p(y), num_columns, _p(x), num_rows * rowsize, rowsize, _p(colidx), _p(a), _p(ad))

    def csr_from_coo_sd(self, num_rows, num_columns, rowidx, colidx, a, symmetric: bool = False):
        nnz = len(a)
        rowptr = np.zeros(num_rows + 1, dtype=np.int64)
        csrsize, lo, hi, ds = _I64(), _I64(), _I64(), _I64()
        f = self.lib.ref_csr_from_coo_sd_size
        f.argtypes = [C.c_int, _I64, _I64, _I64, _P, _P, _P, _P, C.POINTER(_I64), C.POINTER(_I64),
                      C.POINTER(_I64), C.POINTER(_I64)]
        assert f(int(symmetric), num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr),
                 C.byref(csrsize), C.byref(lo), C.byref(hi), C.byref(ds)) == 0
        cc = np.empty(max(csrsize.value, 1), dtype=_idt(self.bits))[:csrsize.value]
        ca = np.empty(max(csrsize.value, 1), dtype=np.float64)[:csrsize.value]
        ad = np.empty(max(num_rows, 1), dtype=np.float64)[:num_rows]
        g = self.lib.ref_csr_from_coo_sd
        g.argtypes = [C.c_int, _I64, _I64, _I64, _P, _P, _P, _P, _I64, _I64, _I64, _P, _P, _P]
        assert g(int(symmetric), num_rows, num_columns, nnz, _p(rowidx), _p(colidx), _p(a), _p(rowptr),
                 csrsize.value, lo.value, hi.value, _p(cc), _p(ca), _p(ad)) == 0
        return rowptr, cc, ca, ad, lo.value, hi.value, ds.value

    def csrgemvsd(self, num_rows, y, num_columns, x, rowptr, colidx, a, ad, rowsizemin=0, rowsizemax=0) -> None:
        f = self.lib.ref_csrgemvsd
        f.argtypes = [_I64, _P, _I64, _P, _I64, _I64, _I64, _P, _P, _P, _P]
        assert f(num_rows, _p(y), num_columns, _p(x), int(rowptr[num_rows]), rowsizemin, rowsizemax, _p(rowptr),
                 _p(colidx), _p(a), _p(ad)) == 0

    def rowsort(self, num_rows, num_columns, rowptr, rowsizemax, colidx, a) -> None:
        f = self.lib.ref_rowsort
        f.argtypes = [_I64, _I64, _P, _I64, _P, _P]
        assert f(num_rows, num_columns, _p(rowptr), rowsizemax, _p(colidx), _p(a)) == 0
