"""ctypes access to the host programs' reader + converters
(ellspmv_b200/host/bin/libhost{32,64}.so) and helpers to write .mtx files."""
import ctypes as C
import os
import subprocess

import numpy as np

from conftest import ROOT

HOST_DIR = os.path.join(ROOT, "ellspmv_b200", "host")
BIN = os.path.join(HOST_DIR, "bin")


def build_host(force=False):
    """Build the library and the host programs unless they are already there
    (the GPU box receives them prebuilt; file times do not survive the copy,
    so `make` would rebuild everything)."""
    lib = os.path.join(ROOT, "ellspmv_b200", "lib", "libellspmv_cuda.so")
    progs = [os.path.join(BIN, p) for p in ("ellspmv", "ellspmv64", "csrspmv", "csrspmv64", "libhost32.so", "libhost64.so")]
    if not force and os.path.exists(lib) and all(os.path.exists(p) for p in progs):
        return
    subprocess.run(["make", "-C", os.path.join(ROOT, "ellspmv_b200", "csrc"), "-j8"], check=True, stdout=subprocess.DEVNULL)
    subprocess.run(["make", "-C", HOST_DIR, "-j8"], check=True, stdout=subprocess.DEVNULL)


def hostlib(bits):
    path = os.path.join(BIN, f"libhost{bits}.so")
    if not os.path.exists(path):
        build_host()
    lib = C.CDLL(path)
    assert lib.host_idx_bits() == bits
    lib.host_free.argtypes = [C.c_void_p]
    lib.host_free.restype = None
    return lib


def _take(lib, ptr, n, ctype, dtype):
    out = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(max(n, 1),))[:n].astype(dtype, copy=True)
    lib.host_free(ptr)
    return out


def ell_from_file(bits, path, gzip=0):
    lib = hostlib(bits)
    dims = (C.c_int64 * 7)()
    colidx, a = C.c_void_p(), C.c_void_p()
    err = lib.host_ell_from_file(path.encode(), gzip, dims, C.byref(colidx), C.byref(a))
    if err:
        return err, list(dims), None, None
    it = C.c_int32 if bits == 32 else C.c_int64
    n = dims[4]
    return 0, list(dims), _take(lib, colidx, n, it, np.int32 if bits == 32 else np.int64), \
        _take(lib, a, n, C.c_double, np.float64)


def csr_from_file(bits, path, gzip=0):
    lib = hostlib(bits)
    dims = (C.c_int64 * 7)()
    rowptr, colidx, a = C.c_void_p(), C.c_void_p(), C.c_void_p()
    err = lib.host_csr_from_file(path.encode(), gzip, dims, C.byref(rowptr), C.byref(colidx), C.byref(a))
    if err:
        return err, list(dims), None, None, None
    it = C.c_int32 if bits == 32 else C.c_int64
    n = dims[3]
    return 0, list(dims), _take(lib, rowptr, dims[0] + 1, C.c_int64, np.int64), \
        _take(lib, colidx, n, it, np.int32 if bits == 32 else np.int64), _take(lib, a, n, C.c_double, np.float64)


def write_mtx(path, nrows, ncols, ri, ci, a, field="real", symmetry="general", comments=("% a comment",)):
    with open(path, "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {symmetry}\n")
        for c in comments:
            f.write(c + "\n")
        f.write(f"{nrows} {ncols} {len(ri)}\n")
        for k in range(len(ri)):
            if field == "pattern":
                f.write(f"{ri[k]} {ci[k]}\n")
            elif field == "integer":
                f.write(f"{ri[k]} {ci[k]} {int(a[k])}\n")
            else:
                f.write(f"{ri[k]} {ci[k]} {float(a[k]):.17g}\n")


def write_vec(path, v):
    with open(path, "w") as f:
        f.write("%%MatrixMarket vector array real general\n")
        f.write(f"{len(v)}\n")
        for t in v:
            f.write(f"{float(t):.17g}\n")
