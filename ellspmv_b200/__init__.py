"""ellspmv_b200 -- Python binding (ctypes) of the B200 SpMV C-ABI library.

The product is the C ABI in ``include/ellspmv_cuda.h`` (built in-tree as
``ellspmv_b200/lib/libellspmv_cuda.so``) and the C host programs under
``ellspmv_b200/host``.  This module only exposes that ABI to Python for the
tests and ``bench.py``; it holds no compute of its own and has no CPU
fallback: importing it without the built library raises, and every compute
call fails with ``ENODEV`` on a machine without a CUDA device.

Names follow the reference (jamtrott/ellspmv): ``ellgemv`` / ``csrgemv`` take
the same arguments, in the same order, as the reference's kernels
(ellspmv.c:1129-1137, csrspmv.c:1565-1575) and return ``0`` or an errno.
"""
from __future__ import annotations

import ctypes as C
import errno
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# ELLSPMV_B200_LIB: another build of the same ABI (A/B timing of two builds on one box); never a fallback
LIB_PATH = os.environ.get("ELLSPMV_B200_LIB") or os.path.join(_HERE, "lib", "libellspmv_cuda.so")

# ---- constants (include/ellspmv_cuda.h) -----------------------------------
KERNEL_AUTO, KERNEL_THREAD, KERNEL_WARP = 0, 1, 2
KERNEL_LONGROW = 4      # ELL: few, long rows (CTA per row group, bit-exact); AUTO picks it by shape
KERNEL_CSR_SCALAR = 3   # CSR only: thread-per-row, bit-exact, for balanced rows (auto picks it)
FMA = 1 << 4
L2_PERSIST_X = 1 << 5
NARROW_INDEX = 1 << 6
COLUMN_BLOCKED = 1 << 7
STAGED_GATHER = 1 << 17
NO_PATTERN = 1 << 18
NO_STAGED_GATHER = 1 << 19
SKIP_PADDING = 1 << 20
PATTERN_MASKS = 1 << 21
NO_PATTERN_LANES = 1 << 23
VALUE_PATTERN = 1 << 24     # opt-in: coefficients from the pattern dictionary too
FUSED_SYNC = 1 << 22
KERNEL_CSR_SELL = 5     # CSR: SELL-128-sigma (AUTO takes it for unbalanced rows)
WIDE_INDEX = 1 << 16
ROWS_PER_THREAD_SHIFT = 8
VARIANT_SHIFT = 12
ACCUMULATE, OVERWRITE, ITERATE = 0, 1, 2
GEN_LAPLACE2D, GEN_STENCIL27, GEN_RANDOM = 1, 2, 3


def rows_per_thread(r: int) -> int:
    return (r & 0x7) << ROWS_PER_THREAD_SHIFT


def variant(v: int) -> int:
    return (v & 0xF) << VARIANT_SHIFT


class EllspmvCudaError(RuntimeError):
    def __init__(self, err: int, where: str, detail: str):
        self.errno = err
        name = errno.errorcode.get(err, str(err))
        super().__init__(f"{where}: {name} ({os.strerror(err)}): {detail}")


class Info(C.Structure):
    _fields_ = [
        ("num_rows", C.c_int64), ("num_columns", C.c_int64), ("rowsize", C.c_int64),
        ("row_begin", C.c_int64), ("global_rows", C.c_int64),
        ("idx_width_bits", C.c_int), ("dev_idx_bits", C.c_int), ("slice_rows", C.c_int),
        ("rows_per_thread", C.c_int), ("kernel", C.c_int), ("fma", C.c_int), ("device", C.c_int),
        ("device_bytes", C.c_int64), ("min_col", C.c_int64), ("max_col", C.c_int64),
        ("launches", C.c_int64), ("num_gpus", C.c_int), ("pattern_rows", C.c_int64),
        ("staged", C.c_int), ("launches_per_spmv", C.c_int), ("tune_ms", C.c_double * 2),
        ("exception_entries", C.c_int64), ("long_rows", C.c_int64), ("sell_slots", C.c_int64),
        ("value_pattern_rows", C.c_int64), ("pattern_id_bytes", C.c_int64),
    ]


class CsrInfo(C.Structure):
    _fields_ = [
        ("num_rows", C.c_int64), ("num_columns", C.c_int64), ("csrsize", C.c_int64),
        ("min_row_len", C.c_int64), ("max_row_len", C.c_int64), ("min_col", C.c_int64), ("max_col", C.c_int64),
        ("device_bytes", C.c_int64), ("kernel", C.c_int), ("sell_slots", C.c_int64), ("sell_real", C.c_int64),
        ("sell_long_rows", C.c_int64), ("sell_long_len", C.c_int64), ("ell_view", C.c_int), ("ell_staged", C.c_int),
        ("launches_per_spmv", C.c_int), ("ell_pattern_rows", C.c_int64), ("num_gpus", C.c_int), ("fma", C.c_int),
        ("ell_pattern_id_bytes", C.c_int64), ("ell_dev_idx_bits", C.c_int), ("ell_rows_per_thread", C.c_int),
    ]


# every symbol include/ellspmv_cuda.h declares: (restype, argtypes)
_P = C.c_void_p
_I64 = C.c_int64
_PROTOTYPES = {
    "ellspmv_cuda_upload": (C.c_int, [C.POINTER(_P), C.c_int, _I64, _I64, _I64, _P, _P, C.c_int, C.c_uint]),
    "ellspmv_cuda_upload_shard": (C.c_int, [C.POINTER(_P), C.c_int, _I64, _I64, _I64, _I64, _I64, _P, _P, C.c_int, C.c_uint]),
    "ellspmv_cuda_upload_coo": (C.c_int, [C.POINTER(_P), C.c_int, _I64, _I64, _I64, _P, _P, _P, C.c_int, C.c_uint]),
    "csrspmv_cuda_upload_coo": (C.c_int, [C.POINTER(_P), C.c_int, _I64, _I64, _I64, _P, _P, _P, C.c_uint]),
    "csrspmv_cuda_download": (C.c_int, [_P, _P, _P, _P]),
    "ellspmv_cuda_generate": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(_I64), C.POINTER(C.c_double), C.c_uint64, C.c_int, _I64, _I64, C.c_int, C.c_uint]),
    "ellspmv_cuda_generate_sharded": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(_I64), C.POINTER(C.c_double), C.c_uint64, C.c_int, C.c_int, C.c_uint]),
    "ellspmv_cuda_spmv": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P]),
    "ellspmv_cuda_spmv_device": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "ellspmv_cuda_spmv_push": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I64), _P]),
    "ellspmv_cuda_spmv_exchange": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, C.POINTER(_P), C.POINTER(_I64), C.POINTER(_I64),
                                             C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(_P), _P, _I64, _P]),
    "ellspmv_cuda_set_diagonal": (C.c_int, [_P, _P, C.c_int]),
    "csrspmv_cuda_set_diagonal": (C.c_int, [_P, _P]),
    "ellspmv_cuda_download": (C.c_int, [_P, _P, _P]),
    "ellspmv_cuda_get_info": (C.c_int, [_P, C.POINTER(Info)]),
    "ellspmv_cuda_free": (None, [_P]),
    "csrspmv_cuda_upload": (C.c_int, [C.POINTER(_P), C.c_int, _I64, _I64, _P, _P, _P, C.c_int, C.c_uint]),
    "csrspmv_cuda_generate": (C.c_int, [C.POINTER(_P), C.c_int, C.POINTER(_I64), C.POINTER(C.c_double), C.c_uint64, C.c_int, C.c_int, C.c_uint]),
    "csrspmv_cuda_spmv": (C.c_int, [_P, _P, _P, C.c_int, C.c_int, _P]),
    "csrspmv_cuda_spmv_device": (C.c_int, [_P, _P, _P, C.c_int, _P]),
    "csrspmv_cuda_device_bytes": (_I64, [_P]),
    "csrspmv_cuda_get_info": (C.c_int, [_P, C.POINTER(CsrInfo)]),
    "csrspmv_cuda_free": (None, [_P]),
    "ellspmv_cuda_malloc_host": (C.c_int, [C.POINTER(_P), _I64]),
    "ellspmv_cuda_free_host": (None, [_P]),
    "ellspmv_cuda_malloc_device": (C.c_int, [C.POINTER(_P), _I64]),
    "ellspmv_cuda_free_device": (None, [_P]),
    "ellspmv_cuda_ipc_export": (C.c_int, [_P, _P]),
    "ellspmv_cuda_ipc_open": (C.c_int, [_P, C.POINTER(_P)]),
    "ellspmv_cuda_ipc_close": (C.c_int, [_P]),
    "ellspmv_cuda_peer_barrier": (C.c_int, [C.c_int, C.c_int, _I64, _P, C.POINTER(_P), _P]),
    "ellspmv_cuda_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "ellspmv_cuda_strerror": (C.c_char_p, [C.c_int]),
    "ellspmv_cuda_last_error": (C.c_char_p, []),
    "ellspmv_cuda_version": (C.c_int, []),
}

_lib: Optional[C.CDLL] = None


def load_library() -> C.CDLL:
    """Load the C-ABI library.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `make -C ellspmv_b200/csrc` "
            "(or python -c 'import __graft_entry__ as g; g.build()'). There is no fallback path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _PROTOTYPES.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library drift apart
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def _check(err: int, where: str) -> None:
    if err != 0:
        detail = load_library().ellspmv_cuda_last_error().decode("utf-8", "replace")
        raise EllspmvCudaError(err, where, detail)


def _ptr(obj) -> Optional[int]:
    """Raw address of a numpy array, a torch tensor (host or device) or an int."""
    if obj is None:
        return None
    if isinstance(obj, int):
        return obj
    if isinstance(obj, np.ndarray):
        if not obj.flags["C_CONTIGUOUS"]:
            raise ValueError("array must be C-contiguous")
        return obj.ctypes.data
    if hasattr(obj, "data_ptr"):
        if not obj.is_contiguous():
            raise ValueError("tensor must be contiguous")
        return obj.data_ptr()
    raise TypeError(f"cannot take the address of {type(obj)!r}")


def _idx_bits(colidx) -> int:
    if isinstance(colidx, np.ndarray):
        if colidx.dtype == np.int32:
            return 32
        if colidx.dtype == np.int64:
            return 64
    elif hasattr(colidx, "dtype"):
        s = str(colidx.dtype)
        if s.endswith("int32"):
            return 32
        if s.endswith("int64"):
            return 64
    raise TypeError("column indices must be int32 or int64")


def device_count() -> int:
    n = C.c_int(0)
    err = load_library().ellspmv_cuda_device_count(C.byref(n))
    return n.value if err == 0 else 0


class EllMatrix:
    """A (shard of a) matrix resident on one GPU in sliced-ELL layout."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    # -- construction -------------------------------------------------------
    @classmethod
    def upload(cls, num_rows: int, num_columns: int, rowsize: int, colidx, a, flags: int = 0,
               *, global_rows: Optional[int] = None, row_begin: int = 0, device: int = -1,
               num_gpus: int = 1) -> "EllMatrix":
        lib = load_library()
        h = C.c_void_p()
        bits = _idx_bits(colidx) if colidx is not None else 32
        if global_rows is None and row_begin == 0 and device < 0:
            err = lib.ellspmv_cuda_upload(C.byref(h), bits, num_rows, num_columns, rowsize,
                                          _ptr(colidx), _ptr(a), num_gpus, flags)
        else:
            g = num_rows if global_rows is None else global_rows
            err = lib.ellspmv_cuda_upload_shard(C.byref(h), bits, g, num_columns, rowsize, row_begin,
                                                row_begin + num_rows, _ptr(colidx), _ptr(a), device, flags)
        _check(err, "ellspmv_cuda_upload")
        return cls(h.value)

    @classmethod
    def upload_coo(cls, num_rows: int, num_columns: int, rowidx, colidx, a, flags: int = 0,
                   device: int = -1) -> "EllMatrix":
        """COO (1-based, file order) -> ELL on the device; same matrix as ell_from_coo + upload."""
        h = C.c_void_p()
        err = load_library().ellspmv_cuda_upload_coo(C.byref(h), _idx_bits(rowidx), num_rows, num_columns, len(a),
                                                     _ptr(rowidx), _ptr(colidx), _ptr(a), device, flags)
        _check(err, "ellspmv_cuda_upload_coo")
        return cls(h.value)

    @classmethod
    def generate(cls, kind: int, dims: Sequence[int], vals: Sequence[float] = (0.0, 0.0), seed: int = 42,
                 idx_bits: int = 32, row_begin: int = 0, row_end: int = -1, device: int = -1,
                 flags: int = 0, num_gpus: int = 1) -> "EllMatrix":
        lib = load_library()
        h = C.c_void_p()
        d = (C.c_int64 * 3)(*(list(dims) + [0, 0, 0])[:3])
        v = (C.c_double * 2)(*vals)
        if num_gpus > 1:
            err = lib.ellspmv_cuda_generate_sharded(C.byref(h), kind, d, v, seed, idx_bits, num_gpus, flags)
            _check(err, "ellspmv_cuda_generate_sharded")
            return cls(h.value)
        err = lib.ellspmv_cuda_generate(C.byref(h), kind, d, v, seed, idx_bits, row_begin, row_end, device, flags)
        _check(err, "ellspmv_cuda_generate")
        return cls(h.value)

    # -- compute --------------------------------------------------------------
    def spmv(self, y: np.ndarray, x: np.ndarray, repeat: int = 1, mode: int = ACCUMULATE) -> np.ndarray:
        """HOST vectors in/out (y is updated in place); returns per-launch seconds."""
        secs = np.zeros(max(repeat, 0), dtype=np.float64)
        err = load_library().ellspmv_cuda_spmv(self._h, _ptr(y), _ptr(x), repeat, mode, _ptr(secs))
        _check(err, "ellspmv_cuda_spmv")
        return secs

    def spmv_device(self, y_dev, x_dev, mode: int = ACCUMULATE, stream: int = 0) -> None:
        err = load_library().ellspmv_cuda_spmv_device(self._h, _ptr(y_dev), _ptr(x_dev), mode, stream or None)
        _check(err, "ellspmv_cuda_spmv_device")

    def spmv_push(self, y_dev, x_dev, mode: int, peer_x: Sequence[int], row_lo: Sequence[int],
                  row_hi: Sequence[int], stream: int = 0) -> None:
        n = len(peer_x)
        px = (C.c_void_p * max(n, 1))(*peer_x)
        lo = (C.c_int64 * max(n, 1))(*row_lo)
        hi = (C.c_int64 * max(n, 1))(*row_hi)
        err = load_library().ellspmv_cuda_spmv_push(self._h, _ptr(y_dev), _ptr(x_dev), mode, n, px, lo, hi,
                                                    stream or None)
        _check(err, "ellspmv_cuda_spmv_push")

    def spmv_exchange(self, y_dev, x_dev, mode: int, peer_x: Sequence[int], row_lo: Sequence[int],
                      row_hi: Sequence[int], rank: int, sync_ranks: Sequence[int], sync_flags: Sequence[int],
                      local_flags: int, epoch: int, stream: int = 0) -> None:
        """Fused SpMV + push + step signalling (ellspmv_cuda_spmv_exchange)."""
        n = len(peer_x)
        px = (C.c_void_p * max(n, 1))(*peer_x)
        lo = (C.c_int64 * max(n, 1))(*row_lo)
        hi = (C.c_int64 * max(n, 1))(*row_hi)
        m = len(sync_ranks)
        sr = (C.c_int * max(m, 1))(*sync_ranks)
        sf = (C.c_void_p * max(m, 1))(*sync_flags)
        err = load_library().ellspmv_cuda_spmv_exchange(self._h, _ptr(y_dev), _ptr(x_dev), mode, n, px, lo, hi,
                                                        rank, m, sr, sf, local_flags, epoch, stream or None)
        _check(err, "ellspmv_cuda_spmv_exchange")

    def set_diagonal(self, ad, order: int = 0) -> None:
        """y <- y + (ad.*x + A*x): the reference's ellgemvsd (order 0) / ellgemv16sd (order 1)."""
        _check(load_library().ellspmv_cuda_set_diagonal(self._h, _ptr(ad), order), "ellspmv_cuda_set_diagonal")

    # -- inspection -------------------------------------------------------------
    def info(self) -> Info:
        out = Info()
        _check(load_library().ellspmv_cuda_get_info(self._h, C.byref(out)), "ellspmv_cuda_get_info")
        return out

    def download(self):
        """Row-major (colidx, a) host arrays, in the caller's index width."""
        i = self.info()
        dt = np.int32 if i.idx_width_bits == 32 else np.int64
        colidx = np.empty(i.num_rows * i.rowsize, dtype=dt)
        a = np.empty(i.num_rows * i.rowsize, dtype=np.float64)
        _check(load_library().ellspmv_cuda_download(self._h, _ptr(colidx), _ptr(a)), "ellspmv_cuda_download")
        return colidx, a

    def free(self) -> None:
        if self._h:
            load_library().ellspmv_cuda_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.free()


class CsrMatrix:
    """A CSR matrix resident on one GPU (comparison path)."""

    def __init__(self, handle: int):
        self._h = C.c_void_p(handle)

    @classmethod
    def upload(cls, num_rows: int, num_columns: int, rowptr, colidx, a, flags: int = 0,
               num_gpus: int = 1) -> "CsrMatrix":
        h = C.c_void_p()
        bits = _idx_bits(colidx) if colidx is not None else 32
        err = load_library().csrspmv_cuda_upload(C.byref(h), bits, num_rows, num_columns, _ptr(rowptr),
                                                 _ptr(colidx), _ptr(a), num_gpus, flags)
        _check(err, "csrspmv_cuda_upload")
        return cls(h.value)

    @classmethod
    def upload_coo(cls, num_rows: int, num_columns: int, rowidx, colidx, a, flags: int = 0) -> "CsrMatrix":
        """COO (1-based, file order) -> CSR on the device; same arrays as csr_from_coo (general matrices)."""
        h = C.c_void_p()
        err = load_library().csrspmv_cuda_upload_coo(C.byref(h), _idx_bits(rowidx), num_rows, num_columns, len(a),
                                                     _ptr(rowidx), _ptr(colidx), _ptr(a), flags)
        _check(err, "csrspmv_cuda_upload_coo")
        m = cls(h.value)
        m._shape = (num_rows, len(a), _idx_bits(rowidx))
        return m

    def download(self, num_rows: int, csrsize: int, idx_bits: int):
        rowptr = np.empty(num_rows + 1, dtype=np.int64)
        colidx = np.empty(max(csrsize, 1), dtype=np.int32 if idx_bits == 32 else np.int64)[:csrsize]
        a = np.empty(max(csrsize, 1), dtype=np.float64)[:csrsize]
        _check(load_library().csrspmv_cuda_download(self._h, _ptr(rowptr), _ptr(colidx), _ptr(a)), "csrspmv_cuda_download")
        return rowptr, colidx, a

    @classmethod
    def generate(cls, kind: int, dims: Sequence[int], seed: int = 42, idx_bits: int = 32, device: int = -1,
                 flags: int = 0, vals: Sequence[float] = (0.0, 0.0)) -> "CsrMatrix":
        """GEN_RANDOM (rows of exactly K entries), or GEN_LAPLACE2D / GEN_STENCIL27 as csr_from_coo
        stores them (no padding; vals = (centre, off-diagonal))."""
        h = C.c_void_p()
        d = (C.c_int64 * 3)(*(list(dims) + [0, 0, 0])[:3])
        v = (C.c_double * 2)(float(vals[0]), float(vals[1]))
        err = load_library().csrspmv_cuda_generate(C.byref(h), kind, d, v, seed, idx_bits, device, flags)
        _check(err, "csrspmv_cuda_generate")
        return cls(h.value)

    def spmv(self, y: np.ndarray, x: np.ndarray, repeat: int = 1, mode: int = ACCUMULATE) -> np.ndarray:
        secs = np.zeros(max(repeat, 0), dtype=np.float64)
        err = load_library().csrspmv_cuda_spmv(self._h, _ptr(y), _ptr(x), repeat, mode, _ptr(secs))
        _check(err, "csrspmv_cuda_spmv")
        return secs

    def spmv_device(self, y_dev, x_dev, mode: int = ACCUMULATE, stream: int = 0) -> None:
        err = load_library().csrspmv_cuda_spmv_device(self._h, _ptr(y_dev), _ptr(x_dev), mode, stream or None)
        _check(err, "csrspmv_cuda_spmv_device")

    def set_diagonal(self, ad) -> None:
        """y <- y + (ad.*x + A*x): the reference's csrgemvsd."""
        _check(load_library().csrspmv_cuda_set_diagonal(self._h, _ptr(ad)), "csrspmv_cuda_set_diagonal")

    def device_bytes(self) -> int:
        return load_library().csrspmv_cuda_device_bytes(self._h)

    def info(self) -> CsrInfo:
        out = CsrInfo()
        _check(load_library().csrspmv_cuda_get_info(self._h, C.byref(out)), "csrspmv_cuda_get_info")
        return out

    def describe(self) -> str:
        """Which kernels a launch of this handle runs."""
        i = self.info()
        arith = "fma (tolerance)" if i.fma else "mul-then-add (bit-exact)"
        if i.ell_view:
            how = "rows of one length" if i.ell_view == 2 else "per-row lengths"
            path = ("staged gather (column blocks), then thread-per-row" if i.ell_staged else "thread-per-row")
            pat = (f", offset patterns on {100.0 * i.ell_pattern_rows / max(i.num_rows, 1):.1f} % of the rows"
                   if i.ell_pattern_rows else "")
            return f"sliced-ELL view of the CSR rows ({how}, width {i.max_row_len}): {path}, {arith}{pat}"
        name = {1: "smem-staged stream kernel", 2: "sub-warp per row + shuffle tree (tolerance)",
                3: "scalar thread-per-row",
                5: f"SELL-128-sigma (rows sorted by length in windows of 4096, width per slice; the {i.sell_long_rows} rows "
                   f"longer than {i.sell_long_len} one CTA each on a second stream)"}
        return f"native CSR: {name.get(i.kernel, str(i.kernel))}, {arith}"

    def free(self) -> None:
        if self._h:
            load_library().csrspmv_cuda_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


# ---- reference-shaped operators -------------------------------------------------
def ellgemv(num_rows, y, num_columns, x, ellsize, rowsize, colidx, a, flags: int = 0) -> int:
    """Drop-in for the reference's ``ellgemv`` (ellspmv.c:1129-1153): y += A*x.

    Same argument order and meaning; returns 0 or an errno value.  One-shot
    convenience (upload, one launch, free) -- the host programs keep the
    handle across the repeat loop instead.
    """
    if ellsize != num_rows * rowsize:
        return errno.EINVAL
    try:
        A = EllMatrix.upload(num_rows, num_columns, rowsize, colidx, a, flags)
        try:
            A.spmv(y, x, 1, ACCUMULATE)
        finally:
            A.free()
    except EllspmvCudaError as e:
        return e.errno
    return 0


def csrgemv(num_rows, y, num_columns, x, csrsize, rowsizemin, rowsizemax, rowptr, colidx, a,
            flags: int = 0) -> int:
    """Drop-in for the reference's ``csrgemv`` (csrspmv.c:1565-1595): y += A*x."""
    del csrsize, rowsizemin, rowsizemax
    try:
        A = CsrMatrix.upload(num_rows, num_columns, rowptr, colidx, a, flags)
        try:
            A.spmv(y, x, 1, ACCUMULATE)
        finally:
            A.free()
    except EllspmvCudaError as e:
        return e.errno
    return 0
