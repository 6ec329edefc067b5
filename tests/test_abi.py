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
    # 5 int64, 7 int, pad, 4 int64, 1 int, pad
    assert C.sizeof(E.Info) == 5 * 8 + 7 * 4 + 4 + 4 * 8 + 8


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
