"""The C-ABI library loads, exports every symbol include/ellspmv_cuda.h
declares, and fails loudly (ENODEV, no CPU fallback) without a GPU."""
import ctypes as C
import errno
import os
import re

import numpy as np
import pytest

import ellspmv_b200 as E
from conftest import ROOT


def declared_functions():
    text = open(os.path.join(ROOT, "include", "ellspmv_cuda.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:ellspmv|csrspmv)_cuda_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in E._PROTOTYPES, f"{n} has no Python prototype"
    assert sorted(E._PROTOTYPES) == names


def test_version_and_strerror(lib):
    assert lib.ellspmv_cuda_version() == 100
    assert lib.ellspmv_cuda_strerror(0) == b"success"
    assert b"nvalid" in lib.ellspmv_cuda_strerror(errno.EINVAL)


def test_info_struct_layout_matches_header():
    # 5 int64, 7 int, pad, 4 int64, 1 int, pad, 1 int64, 2 int, 2 double, 3 int64
    assert C.sizeof(E.Info) == 5 * 8 + 7 * 4 + 4 + 4 * 8 + 8 + 8 + 2 * 4 + 2 * 8 + 5 * 8


def test_product_never_imports_the_oracle():
    for base, _, files in os.walk(os.path.join(ROOT, "ellspmv_b200")):
        for f in files:
            if f.endswith((".py", ".c", ".h", ".cu", ".cuh")):
                src = open(os.path.join(base, f)).read()
                assert "oracle" not in src.lower(), f"{f} mentions the oracle: the product must not use it"


@pytest.mark.skipif(E.device_count() > 0, reason="only meaningful on a machine without a GPU")
def test_no_gpu_means_enodev_not_a_fallback(lib):
    colidx = np.zeros(4, dtype=np.int32)
    a = np.ones(4)
    with pytest.raises(E.EllspmvCudaError) as ei:
        E.EllMatrix.upload(2, 2, 2, colidx, a)
    assert ei.value.errno == errno.ENODEV
    y, x = np.zeros(2), np.ones(2)
    assert E.ellgemv(2, y, 2, x, 4, 2, colidx, a) == errno.ENODEV
    assert np.array_equal(y, np.zeros(2))          # nothing was computed on the CPU
    rowptr = np.array([0, 2, 4], dtype=np.int64)
    assert E.csrgemv(2, y, 2, x, 4, 2, 2, rowptr, colidx, a) == errno.ENODEV
    with pytest.raises(E.EllspmvCudaError):
        E.EllMatrix.generate(E.GEN_LAPLACE2D, (4, 4), (4.0, -1.0))


def test_argument_validation_is_reported_as_einval(lib):
    h = C.c_void_p()
    assert lib.ellspmv_cuda_upload(C.byref(h), 16, 1, 1, 1, None, None, 1, 0) == errno.EINVAL
    assert b"idx_width_bits" in lib.ellspmv_cuda_last_error()
    assert lib.ellspmv_cuda_upload(C.byref(h), 32, -1, 1, 1, None, None, 1, 0) == errno.EINVAL
    assert lib.ellspmv_cuda_spmv(None, None, None, 1, 0, None) == errno.EINVAL
    lib.ellspmv_cuda_free(None)   # no-op, like free(NULL)


def test_header_is_plain_c_and_links(tmp_path, lib):
    """The boundary is a C ABI: the header must compile as C99 and a C caller
    must link against the library with nothing but -lellspmv_cuda."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <errno.h>
#include <stdio.h>
#include "ellspmv_cuda.h"
int main(void) {
    int colidx[4] = {0, 1, 0, 1};
    double a[4] = {1, 2, 3, 4}, x[2] = {1, 1}, y[2] = {0, 0};
    ellspmv_cuda_matrix *A = NULL;
    int n = -1;
    ellspmv_cuda_device_count(&n);
    int err = ellspmv_cuda_upload(&A, 32, 2, 2, 2, colidx, a, 1, 0);
    if (n <= 0) { printf("nodev %d %s\n", err == ENODEV, ellspmv_cuda_strerror(err)); return err == ENODEV ? 0 : 1; }
    if (err) return 2;
    err = ellspmv_cuda_spmv(A, y, x, 1, ELLSPMV_CUDA_ACCUMULATE, NULL);
    ellspmv_cuda_free(A);
    printf("y %g %g\n", y[0], y[1]);
    return (err == 0 && y[0] == 3 && y[1] == 7) ? 0 : 3;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(E.LIB_PATH)
    cmd = ["/usr/bin/gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
           "-L", libdir, "-lellspmv_cuda", f"-Wl,-rpath,{libdir}", "-o", str(exe)]
    subprocess.run(cmd, check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
