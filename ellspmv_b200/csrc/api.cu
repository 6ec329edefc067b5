// api.cu -- the C ABI of include/ellspmv_cuda.h.
//
// Host-side glue only: device memory, streams, events and the copies around
// the kernels in ell_kernels.cu / csr_kernels.cu / layout.cu.  The hot loop
// it stands in for is the reference's repeat loop around ellgemv
// (ellspmv.c:1821-1876) and csrgemv (csrspmv.c:2834-2901).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "handles.cuh"

namespace ellspmv {

static thread_local char g_last_error[512] = "";

void set_last_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int cuda_to_errno(cudaError_t e)
{
    switch (e) {
    case cudaSuccess: return 0;
    case cudaErrorMemoryAllocation: return ENOMEM;
    case cudaErrorInvalidValue:
    case cudaErrorInvalidDevicePointer:
    case cudaErrorInvalidDevice: return EINVAL;
    case cudaErrorNoDevice:
    case cudaErrorInsufficientDriver:
    case cudaErrorInitializationError: return ENODEV;
    case cudaErrorNotSupported: return ENOTSUP;
    default: return EIO;
    }
}

}  // namespace ellspmv

using namespace ellspmv;

namespace {

int check_device(int *device)
{
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count <= 0) {
        cudaGetLastError();
        ELL_FAIL(ENODEV, "no CUDA device available (%s); this library has no CPU fallback",
                 e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (*device < 0) ELL_CK(cudaGetDevice(device));
    if (*device >= count) ELL_FAIL(EINVAL, "device %d out of range (have %d)", *device, count);
    return 0;
}

// Decide layout + kernel from shape and flags.
int configure(ellspmv_cuda_matrix *A, unsigned flags)
{
    A->flags = flags;
    int R = (flags & ELLSPMV_CUDA_ROWS_PER_THREAD_MASK) >> ELLSPMV_CUDA_ROWS_PER_THREAD_SHIFT;
    // auto: enough loads in flight per thread to cover the HBM latency.  With long rows one
    // row per thread does it and keeps the pattern groups small (K = 27: R = 1 / 2 / 4 give
    // 2.09 / 2.57 / 2.71 ms on BASELINE config 3, coverage 83 / 67 / 33 %); with K = 5 a thread
    // that owns one row has 48 bytes in flight and the kernel is latency-bound once the offset
    // patterns (pattern.cu) take the index stream away: 0.649 / 0.603 / 0.608 ms on config 2
    // (profiles/r1_offset_patterns.md).  The boundary at K = 12 is interpolated.
    A->rpt_auto = R == 0;
    if (R == 0) R = A->lay.rowsize <= 12 ? 2 : 1;          // K <= 12: revisited in build_patterns once the coverage is known
    if (R != 1 && R != 2 && R != 4) ELL_FAIL(EINVAL, "rows per thread must be 1, 2 or 4 (got %d)", R);
    int kernel = flags & ELLSPMV_CUDA_KERNEL_MASK;
    A->kernel_auto = kernel == ELLSPMV_CUDA_KERNEL_AUTO;
    if (kernel == ELLSPMV_CUDA_KERNEL_AUTO) {
        // the nnz-per-row switch: thread-per-row keeps rows * 8 loads in flight, which feeds HBM from
        // ~10^5 rows on at any K (the sliced layout is coalesced at any K); with fewer rows the only
        // parallelism left is inside the rows, and from 64 entries per row on a CTA per row group
        // (ell_longrow.cu) wins -- measured crossover in profiles/r2_k_sweep.md
        long long max_rows = 32768, min_k = 64;
        if (const char *env = getenv("ELLSPMV_CUDA_LONGROW_MAX_ROWS")) max_rows = atoll(env);
        if (const char *env = getenv("ELLSPMV_CUDA_LONGROW_MIN_K")) min_k = atoll(env);
        const bool lr = A->lay.rowsize >= min_k && A->lay.num_rows <= max_rows &&
                        !(flags & (ELLSPMV_CUDA_STAGED_GATHER | ELLSPMV_CUDA_COLUMN_BLOCKED | ELLSPMV_CUDA_VARIANT_MASK |
                                   ELLSPMV_CUDA_ROWS_PER_THREAD_MASK));
        kernel = lr ? kKernelLongRow : ELLSPMV_CUDA_KERNEL_THREAD;
    }
    if (kernel != ELLSPMV_CUDA_KERNEL_THREAD && kernel != ELLSPMV_CUDA_KERNEL_WARP && kernel != kKernelLongRow)
        ELL_FAIL(EINVAL, "unknown kernel selector %d", kernel);
    if (kernel == kKernelLongRow) R = 1;
    A->dev_idx_bits = A->host_idx_bits;
    if (A->host_idx_bits == 64 && !(flags & ELLSPMV_CUDA_WIDE_INDEX) && A->num_columns < (1LL << 31))
        A->dev_idx_bits = 32;     // index narrowing: a pure device-layout choice (bit-exact results)
    A->lay.slice_rows = kernel == kKernelLongRow ? 1 : kBlockThreads * R;    // long rows: the reference's row-major layout
    A->lay.num_slices = (A->lay.num_rows + A->lay.slice_rows - 1) / A->lay.slice_rows;
    cudaDeviceProp prop;
    ELL_CK(cudaGetDeviceProperties(&prop, A->device));
    A->cfg.idx_bits = A->dev_idx_bits;
    A->cfg.rows_per_thread = R;
    A->cfg.kernel = kernel;
    A->cfg.variant = (flags & ELLSPMV_CUDA_VARIANT_MASK) >> ELLSPMV_CUDA_VARIANT_SHIFT;
    A->cfg.fma = (flags & ELLSPMV_CUDA_FMA) != 0;
    A->cfg.num_sms = prop.multiProcessorCount;
    A->cfg.x_bytes = A->num_columns * 8;
    A->cfg.persist_x = false;
    if (flags & ELLSPMV_CUDA_L2_PERSIST_X) {
        size_t want = (size_t)A->cfg.x_bytes;
        size_t maxp = (size_t)prop.persistingL2CacheMaxSize;
        size_t maxw = (size_t)prop.accessPolicyMaxWindowSize;
        if (maxp > 0 && want <= maxw) {
            ELL_CK(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want < maxp ? want : maxp));
            A->cfg.persist_x = true;
        }
    }
    return 0;
}

int alloc_matrix(ellspmv_cuda_matrix *A)
{
    const int64_t n = A->lay.entries();
    const size_t vb = (size_t)(n > 0 ? n : 1) * 8;
    const size_t cb = (size_t)(n > 0 ? n : 1) * (A->dev_idx_bits / 8);
    ELL_CK(cudaMalloc(&A->vals, vb));
    ELL_CK(cudaMalloc(&A->cols, cb));
    ELL_CK(cudaMalloc(&A->d_minmax, 2 * sizeof(long long)));
    ELL_CK(cudaStreamCreateWithFlags(&A->stream, cudaStreamNonBlocking));
    A->device_bytes = (int64_t)(vb + cb);
    // the tail of the last slice must hold harmless entries (col 0, 0.0)
    ELL_CK(cudaMemsetAsync(A->vals, 0, vb, A->stream));
    ELL_CK(cudaMemsetAsync(A->cols, 0, cb, A->stream));
    ELL_CK(init_minmax(A->d_minmax, A->stream));
    return 0;
}

// First use of a kernel pays for loading its code; do that at upload time so the
// first timed launch is a steady-state one (an empty launch: zero rows).
void warm_kernels(ellspmv_cuda_matrix *A)
{
    preload_sync_kernels();
    if (A->lay.rowsize <= 0) { cudaGetLastError(); return; }
    EllSpmvArgs args = {};
    args.vals = A->vals; args.cols = A->cols; args.x = A->vals; args.y = A->vals;
    args.num_rows = 0;
    args.rowsize = A->lay.rowsize;
    args.beta = 1;
    args.patid = A->pat.max_explicit ? nullptr : A->pat.patid;
    args.patinfo = A->pat.max_explicit ? A->pat.patinfo : nullptr;
    args.patlane = A->pat.patlane;
    args.pat = A->pat.pat;
    args.vpat = A->pat.vpat;
    args.rowlen = A->d_rowlen;
    // every instantiation a later launch of this handle may pick: y aligned for vector access or
    // not (a shard that starts at an odd row), with and without the in-kernel hand-shake.  A first
    // launch loads code, and a load can wait for a kernel that is spinning for this one's signal.
    long long dummy_flags[1] = {0};
    for (int misaligned = 0; misaligned < 2; misaligned++)
        for (int synced = 0; synced < 2; synced++) {
            if (synced && !(A->flags & ELLSPMV_CUDA_FUSED_SYNC)) continue;
            args.y = reinterpret_cast<double *>(reinterpret_cast<char *>(A->vals) + (misaligned ? 8 : 0));
            args.sync = StepSync{};
            if (synced) { args.sync.local_flags = dummy_flags; args.sync.num_ranges = 0; }   // never dereferenced: no rows
            if (launch_ell_spmv(A->cfg, args, 1, A->stream) != cudaSuccess) cudaGetLastError();
        }
    cudaStreamSynchronize(A->stream);
    cudaGetLastError();
}

int finish_minmax(ellspmv_cuda_matrix *A)
{
    long long mm[2];
    ELL_CK(cudaMemcpyAsync(mm, A->d_minmax, sizeof(mm), cudaMemcpyDeviceToHost, A->stream));
    ELL_CK(cudaStreamSynchronize(A->stream));
    A->min_col = mm[0];
    A->max_col = mm[1];
    warm_kernels(A);
    // the max accumulator starts at -1 and negative indices never raise it, so the lower
    // bound is checked on its own (min still at its start value = no entries at all)
    const bool any = mm[0] != 0x7fffffffffffffffLL;
    if (any && (mm[0] < 0 || mm[1] >= A->num_columns))
        ELL_FAIL(EINVAL, "column index out of range: [%lld, %lld] with %lld columns",
                 mm[0], mm[1] < mm[0] ? mm[0] : mm[1], (long long)A->num_columns);
    if (!any) { A->min_col = 0; A->max_col = -1; }
    return 0;
}

// Re-lay a handle's matrix for another number of rows per thread (slice height 128 * R): a pure
// permutation on the device, through two row-major staging buffers.  ENOMEM (the second copy does
// not fit next to the first) leaves the handle untouched.
int change_rows_per_thread(ellspmv_cuda_matrix *A, int R)
{
    EllLayout nl = A->lay;
    nl.slice_rows = kBlockThreads * R;
    nl.num_slices = (nl.num_rows + nl.slice_rows - 1) / nl.slice_rows;
    const int64_t rows = A->lay.num_rows, K = A->lay.rowsize, n = nl.entries();
    const int ib = A->dev_idx_bits / 8;
    double *nv = nullptr, *sv = nullptr;
    void *nc = nullptr, *sc = nullptr;
    int64_t chunk_rows = (64LL << 20) / (K * (8 + ib));
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > rows) chunk_rows = rows;
    auto cleanup = [&]() { cudaFree(nv); cudaFree(nc); cudaFree(sv); cudaFree(sc); };
    cudaError_t ce = cudaMalloc(&nv, (size_t)n * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&nc, (size_t)n * ib);
    if (ce == cudaSuccess) ce = cudaMalloc(&sv, (size_t)chunk_rows * K * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&sc, (size_t)chunk_rows * K * ib);
    if (ce == cudaErrorMemoryAllocation) { cudaGetLastError(); cleanup(); return ENOMEM; }
    // the tail of the last slice must hold harmless entries (col 0, 0.0)
    if (ce == cudaSuccess) ce = cudaMemsetAsync(nv, 0, (size_t)n * 8, A->stream);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(nc, 0, (size_t)n * ib, A->stream);
    for (int64_t r0 = 0; r0 < rows && ce == cudaSuccess; r0 += chunk_rows) {
        const int64_t m = (rows - r0 < chunk_rows) ? rows - r0 : chunk_rows;
        ce = unlayout_chunk(A->dev_idx_bits, A->dev_idx_bits, A->cols, A->vals, sc, sv, A->lay, r0, m, A->stream);
        if (ce == cudaSuccess)
            ce = relayout_chunk(A->dev_idx_bits, A->dev_idx_bits, sc, sv, nc, nv, nl, r0, m, A->d_minmax, A->stream);
    }
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(A->stream);
    if (ce != cudaSuccess) {
        cleanup();
        set_last_error("re-layout: %s", cudaGetErrorString(ce));
        return cuda_to_errno(ce);
    }
    cudaFree(sv); cudaFree(sc);
    A->device_bytes += (n - A->lay.entries()) * (8 + ib);
    cudaFree(A->vals); cudaFree(A->cols);
    A->vals = nv; A->cols = nc;
    A->lay = nl;
    A->cfg.rows_per_thread = R;
    return 0;
}

// Offset patterns (pattern.cu): groups of 32 rows whose column indices are row + d[l] stop
// reading the index stream.  On by default for the thread-per-row kernel;
// ELLSPMV_CUDA_NO_PATTERN turns it off.
int build_patterns(ellspmv_cuda_matrix *A)
{
    if ((A->flags & ELLSPMV_CUDA_NO_PATTERN) || A->cfg.kernel != ELLSPMV_CUDA_KERNEL_THREAD ||
        A->lay.num_rows <= 0 || A->lay.rowsize <= 0)
        return 0;
    // value patterns (opt-in; a constant-coefficient stencil then streams neither indices nor values): bit-exact mode only
    const double *pv = (!(A->flags & ELLSPMV_CUDA_VALUE_PATTERN) || A->cfg.fma) ? nullptr : A->vals;
    cudaError_t ce = pattern_build(&A->pat, A->dev_idx_bits, A->cols, pv, A->lay, A->cfg.rows_per_thread, A->row_begin,
                                   (A->flags & ELLSPMV_CUDA_PATTERN_MASKS) ? 4 : 0,
                                   !(A->flags & ELLSPMV_CUDA_NO_PATTERN_LANES), A->stream);
    if (ce != cudaSuccess) { set_last_error("offset patterns: %s", cudaGetErrorString(ce)); return cuda_to_errno(ce); }
    // AUTO rows per thread, second look.  Two rows per thread (K <= 12) pay only while the 64-row groups
    // keep their patterns: a 3D grid puts a boundary row into most groups of 64 (7-point 224^3: 57 % of
    // the rows patterned at two rows per thread, 100 % at one -- 0.195 vs 0.168 ms,
    // profiles/r2_stencil_sweep.md).  Such a matrix is re-laid with one row per thread and searched again;
    // a matrix without any pattern stays as it is (both layouts then run at the same speed).
    if (A->rpt_auto && A->cfg.rows_per_thread == 2 && A->pat.any() && A->pat.covered * 10 < A->pat.groups * 9) {
        const int e2 = change_rows_per_thread(A, 1);
        if (e2 == 0) {
            pattern_free(&A->pat);
            ce = pattern_build(&A->pat, A->dev_idx_bits, A->cols, pv, A->lay, A->cfg.rows_per_thread, A->row_begin,
                               (A->flags & ELLSPMV_CUDA_PATTERN_MASKS) ? 4 : 0,
                               !(A->flags & ELLSPMV_CUDA_NO_PATTERN_LANES), A->stream);
            if (ce != cudaSuccess) { set_last_error("offset patterns: %s", cudaGetErrorString(ce)); return cuda_to_errno(ce); }
        } else if (e2 != ENOMEM) {
            return e2;
        }
    }
    A->device_bytes += A->pat.bytes;
    if (A->pat.any()) warm_kernels(A);
    return 0;
}

// ELLSPMV_CUDA_COLUMN_BLOCKED: bin the entries by column block so that each
// block's slice of x stays in L2 (ell_blocked.cu); no-op when x already fits
int launch(ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta,
           const PushTargets *push, cudaStream_t stream, int64_t slice_begin, int64_t num_slices,
           const StepSync *sync);

// KERNEL_AUTO for matrices with scattered columns (BASELINE config 4): when x does not fit in
// L2, no offset pattern was found and a warp's gather touches many different lines, build the
// staged gather, time it against the direct gather on scratch vectors and keep the faster one.
// Both give the same bits (ell_staged.cu), so the choice is invisible in the results.
int auto_staged_gather(ellspmv_cuda_matrix *A, long long block_bytes)
{
    const int64_t x_bytes = A->num_columns * 8;
    if (!A->kernel_auto || (A->flags & (ELLSPMV_CUDA_NO_STAGED_GATHER | ELLSPMV_CUDA_COLUMN_BLOCKED)) ||
        A->cfg.kernel != ELLSPMV_CUDA_KERNEL_THREAD || (A->cfg.variant & 1))
        return 0;
    long long min_x = 96LL << 20;                    // L2 is 126 MB: below this x mostly stays resident by itself
    if (const char *env = getenv("ELLSPMV_CUDA_AUTO_STAGED_MIN_X_BYTES")) min_x = atoll(env);
    if (x_bytes <= min_x || A->pat.any()) return 0;
    double lines = 0.0;
    ELL_CK(sg_scatter_estimate(A->dev_idx_bits, A->cols, A->lay, &lines, A->stream));
    if (lines < 16.0) return 0;                      // gathers mostly share lines: L1/L2 serve them
    size_t free_b = 0, total_b = 0;
    ELL_CK(cudaMemGetInfo(&free_b, &total_b));
    const int64_t scratch = (A->num_columns + A->lay.num_rows) * 8;
    if ((int64_t)free_b < sg_bytes_estimate(A->dev_idx_bits, A->lay) + scratch + (1LL << 30)) return 0;
    cudaError_t ce = sg_build(&A->sg, A->dev_idx_bits, A->cols, A->vals, A->lay, A->num_columns, block_bytes, A->stream);
    if (ce == cudaErrorMemoryAllocation) { cudaGetLastError(); A->sg = nullptr; return 0; }
    if (ce != cudaSuccess) { set_last_error("staged gather: %s", cudaGetErrorString(ce)); return cuda_to_errno(ce); }
    if (!A->sg) return 0;
    // the trial: one warm-up and two timed launches of each path on zeroed scratch vectors
    double *sx = nullptr, *sy = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    SgMatrix *sg = A->sg;
    auto cleanup = [&]() { cudaFree(sx); cudaFree(sy); if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
    ce = cudaMalloc(&sx, (size_t)A->num_columns * 8);
    if (ce == cudaSuccess) ce = cudaMalloc(&sy, (size_t)(A->lay.num_rows > 0 ? A->lay.num_rows : 1) * 8);
    if (ce == cudaSuccess) ce = cudaMemsetAsync(sx, 0, (size_t)A->num_columns * 8, A->stream);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e0);
    if (ce == cudaSuccess) ce = cudaEventCreate(&e1);
    if (ce != cudaSuccess) {
        // no room for the trial: keep the staged path only on request
        cudaGetLastError(); cleanup(); sg_free(A->sg); A->sg = nullptr;
        return 0;
    }
    const int64_t launches_before = A->launches;
    int err = 0;
    for (int path = 0; path < 2 && !err; path++) {
        A->sg = path ? sg : nullptr;
        err = launch(A, sy, sx, 0, nullptr, A->stream, 0, -1, nullptr);
        if (!err && cudaEventRecord(e0, A->stream) != cudaSuccess) err = EIO;
        for (int i = 0; i < 2 && !err; i++) err = launch(A, sy, sx, 0, nullptr, A->stream, 0, -1, nullptr);
        if (!err && (cudaEventRecord(e1, A->stream) != cudaSuccess || cudaEventSynchronize(e1) != cudaSuccess)) err = EIO;
        float ms = 0.f;
        if (!err && cudaEventElapsedTime(&ms, e0, e1) != cudaSuccess) err = EIO;
        A->tune_ms[path] = ms / 2.0;
    }
    A->launches = launches_before;
    A->sg = sg;
    cleanup();
    if (err) { set_last_error("staged gather trial failed"); return err; }
    if (A->tune_ms[1] < A->tune_ms[0]) {
        A->staged_mode = 2;
        A->device_bytes += sg_bytes(A->sg);
    } else {
        sg_free(A->sg);
        A->sg = nullptr;
    }
    return 0;
}

int build_column_blocks(ellspmv_cuda_matrix *A)
{
    if (A->lay.num_rows <= 0 || A->lay.rowsize <= 0 || A->cfg.kernel == kKernelLongRow) return 0;
    if (A->flags & ELLSPMV_CUDA_SKIP_PADDING) {
        // SELL-128-sigma copy without the rows' trailing padding (sell.cu); the regular layout stays
        // (download, push, diagonal order 1 and the chunked host call use it)
        cudaError_t ce = sell_build_ell(&A->sell, A->dev_idx_bits, A->cols, A->vals, A->lay, A->row_begin, A->num_columns,
                                        A->stream);
        if (ce != cudaSuccess) { set_last_error("skip padding: %s", cudaGetErrorString(ce)); return cuda_to_errno(ce); }
        A->device_bytes += sell_bytes(A->sell);
        return 0;
    }
    if (!(A->flags & (ELLSPMV_CUDA_COLUMN_BLOCKED | ELLSPMV_CUDA_STAGED_GATHER))) {
        long long target = 48LL << 20;
        if (const char *env = getenv("ELLSPMV_CUDA_BLOCK_BYTES")) {
            long long v = atoll(env);
            if (v >= 8) target = v;
        }
        return auto_staged_gather(A, target);
    }
    // x bytes per column block: 48 MB stays resident in the 126 MB (2 x 63 MB) L2 next to the
    // streaming matrix -- measured 11.6 / 9.2 / 9.8 / 13.5 ms at 32 / 48 / 64 / 80 MB on BASELINE
    // config 4 (profiles/r1_c4_column_blocked.md); ELLSPMV_CUDA_BLOCK_BYTES overrides it
    long long target = 48LL << 20;
    if (const char *env = getenv("ELLSPMV_CUDA_BLOCK_BYTES")) {
        long long v = atoll(env);
        if (v >= 8) target = v;
    }
    if (A->flags & ELLSPMV_CUDA_STAGED_GATHER) {
        // the bit-exact flavour (ell_staged.cu); wins over COLUMN_BLOCKED when both are set
        cudaError_t ce = sg_build(&A->sg, A->dev_idx_bits, A->cols, A->vals, A->lay, A->num_columns, target, A->stream);
        if (ce != cudaSuccess) { set_last_error("staged gather: %s", cudaGetErrorString(ce)); return cuda_to_errno(ce); }
        A->device_bytes += sg_bytes(A->sg);
        if (A->sg) A->staged_mode = 1;
        return 0;
    }
    cudaError_t ce = cb_build(&A->cb, A->dev_idx_bits, A->vals, A->cols, A->lay, A->num_columns, target, A->stream);
    if (ce != cudaSuccess) { set_last_error("column blocking: %s", cudaGetErrorString(ce)); return cuda_to_errno(ce); }
    A->device_bytes += cb_bytes(A->cb);
    return 0;
}

int ensure_vectors(ellspmv_cuda_matrix *A)
{
    int64_t need = A->lay.num_rows > A->num_columns ? A->lay.num_rows : A->num_columns;
    if (need < 1) need = 1;
    if (A->d_x && A->vec_len >= need) return 0;
    if (A->d_x) cudaFree(A->d_x);
    if (A->d_y) cudaFree(A->d_y);
    A->d_x = A->d_y = nullptr;
    ELL_CK(cudaMalloc(&A->d_x, (size_t)need * 8));
    ELL_CK(cudaMalloc(&A->d_y, (size_t)need * 8));
    A->vec_len = need;
    A->device_bytes += 2 * need * 8;
    return 0;
}

// The part of x a handle's kernels read: the column range its stored entries reference
// ([min_col, max_col], from the upload-time reduction) plus, with a separately stored
// diagonal, its own global rows.  The host-vector calls upload only this range: a row shard
// of a stencil needs its own slice and two halo planes, not all of x (the reference passes
// the whole x to every thread, ellspmv.c:1841-1842; over PCIe that would be N copies of it).
void x_range(const ellspmv_cuda_matrix *A, int64_t *lo, int64_t *hi)
{
    int64_t a = A->min_col, b = A->max_col + 1;
    if (b <= a) { a = 0; b = 0; }
    if (A->d_ad && A->lay.num_rows > 0) {
        const int64_t r0 = A->row_begin, r1 = A->row_begin + A->lay.num_rows;
        if (b <= a) { a = r0; b = r1; }
        else { a = r0 < a ? r0 : a; b = r1 > b ? r1 : b; }
    }
    if (b > A->num_columns) b = A->num_columns;
    *lo = a; *hi = b;
}

int ensure_events(std::vector<cudaEvent_t> &ev, size_t n)
{
    while (ev.size() < n) {
        cudaEvent_t e;
        ELL_CK(cudaEventCreate(&e));
        ev.push_back(e);
    }
    return 0;
}

int launch(ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta,
           const PushTargets *push, cudaStream_t stream, int64_t slice_begin = 0, int64_t num_slices = -1,
           const StepSync *sync = nullptr);

// can this handle's SpMV launch carry the fused step synchronisation (ell_thread_kernel only)?
bool fused_sync_capable(const ellspmv_cuda_matrix *A)
{
    return (A->flags & ELLSPMV_CUDA_FUSED_SYNC) && A->cfg.kernel == ELLSPMV_CUDA_KERNEL_THREAD && !(A->cfg.variant & 1) && !A->sg && !A->cb && !A->sell &&
           !A->d_rowlen && !A->pat.max_explicit && !A->pat.patlane && !A->pat.vpat && A->lay.rowsize > 0 && A->lay.num_rows > 0;
}

int launch(ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta,
           const PushTargets *push, cudaStream_t stream, int64_t slice_begin, int64_t num_slices,
           const StepSync *sync)
{
    EllSpmvArgs args = {};
    args.vals = A->vals;
    args.cols = A->cols;
    args.x = x_dev;
    args.y = y_dev;
    args.num_rows = A->lay.num_rows;
    args.row_begin = A->row_begin;
    args.rowsize = A->lay.rowsize;
    args.beta = beta;
    args.slice_begin = slice_begin;
    args.ad = A->d_ad;
    args.sd_order = A->sd_order;
    args.patid = A->pat.max_explicit ? nullptr : A->pat.patid;
    args.patinfo = A->pat.max_explicit ? A->pat.patinfo : nullptr;
    args.patlane = A->pat.patlane;
    args.pat = A->pat.pat;
    args.vpat = A->pat.vpat;
    args.rowlen = A->d_rowlen;
    // value-stream L2 prefetch 128 slices ahead when most rows are patterned (ell_kernels.cu)
    // (long rows have the loads in flight anyway; keep one request at or below 256 KB)
    args.prefetch = (A->pat.any() && !A->pat.vpat && A->pat.covered * 2 >= A->pat.groups &&
                     (int64_t)A->lay.slice_rows * A->lay.rowsize * 8 <= (1 << 18)) ? 128 : 0;
    if (num_slices < 0) num_slices = A->lay.num_slices - slice_begin;
    if (push) args.push = *push; else args.push.num_peers = 0;
    if (sync) args.sync = *sync;          // only passed for a full launch of a fused_sync_capable handle
    if (A->sg && slice_begin == 0 && num_slices == A->lay.num_slices) {
        ELL_CK(sg_spmv(A->sg, x_dev, y_dev, A->d_ad, A->sd_order, A->lay.num_rows,
                       A->row_begin, beta, push, stream, A->d_rowlen));
        A->launches += sg_launches(A->sg);
        return 0;
    }
    if (A->sell && !push && !sync && !(A->d_ad && A->sd_order) && slice_begin == 0 && num_slices == A->lay.num_slices) {
        ELL_CK(sell_spmv(A->sell, A->cfg.fma, nullptr, nullptr, nullptr, 0, x_dev, y_dev, A->d_ad, A->row_begin, beta, stream));
        A->launches += sell_launches(A->sell);
        return 0;
    }
    if (A->cb && !push && !A->d_ad && slice_begin == 0 && num_slices == A->lay.num_slices) {
        ELL_CK(cb_spmv(A->cb, A->cfg.fma, x_dev, y_dev, A->lay.num_rows, A->num_columns, beta, stream));
        A->launches += cb_blocks(A->cb);
        return 0;
    }
    if (A->lay.rowsize == 0 && !A->d_ad) {
        // K = 0: y += 0 for beta=1, y = 0 for beta=0
        if (!beta && A->lay.num_rows > 0)
            ELL_CK(cudaMemsetAsync(y_dev, 0, (size_t)A->lay.num_rows * 8, stream));
        return 0;
    }
    ELL_CK(launch_ell_spmv(A->cfg, args, num_slices, stream));
    A->launches++;
    return 0;
}

// One y <- beta*y + A*x with host vectors, software-pipelined over row chunks:
// three streams, one per engine.  Upload stream: for each chunk the x piece it
// needs and its y rows (beta = 1) go up, back to back -- the host->device link is
// what bounds the call and never waits for a kernel.  Compute stream: the chunk's
// slices run as soon as its upload has landed.  Download stream: the finished y
// rows come back, so the upload of chunk c+1 overlaps the download of chunk c on
// the full-duplex PCIe link.  seconds = sum of the chunk kernels' device-event
// times.
//
// Chunk sizes (profiles/r2_e2e_chunks.md): the call is the upload (x range + y) plus a fill (the
// first chunk's upload, before anything runs) and a drain (the last chunk's download); every copy
// costs ~13 us of set-up on top of its bytes (8 / 16 / 32 / 64 / 128 equal chunks: 21.98 / 21.91 /
// 22.55 / 24.27 / 25.94 ms on BASELINE config 2).  So the rows are cut into 64 units and the chunks
// ramp up and down: 1, 1, 2, 4, then 8 units each, then 4, 2, 1, 1 -- small at both ends (fill and
// drain are 1/64 of the transfer), few copies in between.
int spmv_pipelined(ellspmv_cuda_matrix *A, double *y, const double *x, int beta, double *seconds)
{
    const int64_t rows = A->lay.num_rows, S = A->lay.slice_rows;
    const int64_t slices = A->lay.num_slices;
    std::vector<int> plan = {1, 1, 2, 4, 8, 8, 8, 8, 8, 8, 4, 2, 1, 1};      // chunk sizes in units
    int64_t nunits = 64;
    if (const char *env = getenv("ELLSPMV_CUDA_HOST_CHUNKS")) {              // experiments: that many equal chunks
        const long long v = atoll(env);
        if (v >= 1 && v <= 4096) { nunits = v; plan.assign((size_t)v, 1); }
    }
    if (const char *env = getenv("ELLSPMV_CUDA_HOST_PLAN")) {                // experiments: "1,1,2,4,..." chunk sizes in units
        std::vector<int> p;
        int64_t sum = 0;
        for (const char *q = env; *q;) {
            char *end = nullptr;
            const long v = strtol(q, &end, 10);
            if (end == q || v < 1 || v > 4096) { p.clear(); break; }
            p.push_back((int)v); sum += v;
            q = (*end == ',') ? end + 1 : end;
            if (*end && *end != ',') { p.clear(); break; }
        }
        if (!p.empty() && sum <= 65536) { plan.swap(p); nunits = sum; }
    }
    if (slices < nunits) { nunits = slices; plan.assign((size_t)nunits, 1); }
    const int nchunks = (int)plan.size();
    const int64_t unit_slices = (slices + nunits - 1) / nunits;
    int err = ensure_events(A->events, 3 * (size_t)nchunks + 1);
    if (err) return err;
    if (!A->stream_out) ELL_CK(cudaStreamCreateWithFlags(&A->stream_out, cudaStreamNonBlocking));
    if (!A->stream_in) ELL_CK(cudaStreamCreateWithFlags(&A->stream_in, cudaStreamNonBlocking));
    cudaStream_t s = A->stream, so = A->stream_out, si = A->stream_in;
    int64_t xlo, xhi;
    x_range(A, &xlo, &xhi);
    // x goes up in pieces too: before a chunk runs, x is on the device up to the largest column
    // the chunk references (one-off reduction per handle and unit size).  For a banded or stencil
    // matrix the pieces interleave with the y chunks, so the first kernel starts after 1/64 of the
    // upload instead of after all of x; for a scattered matrix the first chunk needs everything and
    // this degenerates to "x first".
    if ((int64_t)A->chunk_max.size() != nunits) {
        long long *d_cm = nullptr;
        ELL_CK(cudaMalloc(&d_cm, (size_t)nunits * 8));
        std::vector<long long> cm((size_t)nunits, -1);
        cudaError_t ce = chunk_max_cols(A->dev_idx_bits, A->cols, A->lay, unit_slices, (int)nunits, d_cm, s);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(cm.data(), d_cm, (size_t)nunits * 8, cudaMemcpyDeviceToHost, s);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(s);
        cudaFree(d_cm);
        if (ce != cudaSuccess) ELL_FAIL(cuda_to_errno(ce), "chunk column ranges: %s", cudaGetErrorString(ce));
        A->chunk_max.swap(cm);
    }
    int64_t up_hi = xlo;                       // x[xlo, up_hi) is on the device (or on its way, in stream order)
    int used = 0;
    // the uploads overwrite the handle's vectors: order them after whatever the compute stream still holds
    ELL_CK(cudaEventRecord(A->events[3 * (size_t)nchunks], s));
    ELL_CK(cudaStreamWaitEvent(si, A->events[3 * (size_t)nchunks], 0));
    int64_t u0 = 0;
    for (int c = 0; c < nchunks; c++) {
        const int64_t nu = plan[(size_t)c];
        const int64_t s0 = u0 * unit_slices;
        if (s0 >= slices) break;
        const int64_t ns = (slices - s0 < nu * unit_slices) ? slices - s0 : nu * unit_slices;
        const int64_t r0 = s0 * S;
        const int64_t r1 = (r0 + ns * S < rows) ? r0 + ns * S : rows;
        int64_t need = 0;
        for (int64_t u = u0; u < u0 + nu && u < nunits; u++)
            if (A->chunk_max[(size_t)u] + 1 > need) need = A->chunk_max[(size_t)u] + 1;
        if (A->d_ad && A->row_begin + r1 > need) need = A->row_begin + r1;      // ad[i] * x[global row i]
        if (s0 + ns >= slices || need > xhi) need = xhi;
        if (need > up_hi) {
            ELL_CK(cudaMemcpyAsync(A->d_x + up_hi, x + up_hi, (size_t)(need - up_hi) * 8, cudaMemcpyDefault, si));
            up_hi = need;
        }
        if (beta) ELL_CK(cudaMemcpyAsync(A->d_y + r0, y + r0, (size_t)(r1 - r0) * 8, cudaMemcpyDefault, si));
        ELL_CK(cudaEventRecord(A->events[2 * (size_t)nchunks + c], si));
        ELL_CK(cudaStreamWaitEvent(s, A->events[2 * (size_t)nchunks + c], 0));
        ELL_CK(cudaEventRecord(A->events[2 * c], s));
        if ((err = launch(A, A->d_y, A->d_x, beta, nullptr, s, s0, ns))) return err;
        ELL_CK(cudaEventRecord(A->events[2 * c + 1], s));
        ELL_CK(cudaStreamWaitEvent(so, A->events[2 * c + 1], 0));
        ELL_CK(cudaMemcpyAsync(y + r0, A->d_y + r0, (size_t)(r1 - r0) * 8, cudaMemcpyDefault, so));
        used = c + 1;
        u0 += nu;
    }
    ELL_CK(cudaStreamSynchronize(si));
    ELL_CK(cudaStreamSynchronize(s));
    ELL_CK(cudaStreamSynchronize(so));
    if (seconds) {
        double total = 0.0;
        for (int c = 0; c < used; c++) {
            float ms = 0.f;
            ELL_CK(cudaEventElapsedTime(&ms, A->events[2 * c], A->events[2 * c + 1]));
            total += (double)ms * 1e-3;
        }
        seconds[0] = total;
    }
    return 0;
}

// one launch on many rows through kernels that can run a range of slices: worth pipelining
bool pipelined_host_call_applies(const ellspmv_cuda_matrix *A)
{
    return A->lay.num_rows >= (1 << 20) && A->lay.rowsize > 0 && A->num_columns > 0 && !A->cb && !A->sg && !A->sell &&
           A->cfg.kernel != kKernelLongRow;
}

int new_handle(ellspmv_cuda_matrix **out, int idx_width_bits, int64_t global_rows,
               int64_t num_columns, int64_t rowsize, int64_t row_begin, int64_t row_end,
               int device, unsigned flags)
{
    if (!out) ELL_FAIL(EINVAL, "out is NULL");
    *out = nullptr;
    if (idx_width_bits != 32 && idx_width_bits != 64)
        ELL_FAIL(EINVAL, "idx_width_bits must be 32 or 64 (got %d)", idx_width_bits);
    if (global_rows < 0 || num_columns < 0 || rowsize < 0 || rowsize > 0x7fffffff)
        ELL_FAIL(EINVAL, "negative or oversized dimension");
    if (row_begin < 0 || row_end < row_begin || row_end > global_rows)
        ELL_FAIL(EINVAL, "row range [%lld, %lld) outside [0, %lld)", (long long)row_begin,
                 (long long)row_end, (long long)global_rows);
    if (idx_width_bits == 32 && (num_columns > 0x7fffffffLL || global_rows > 0x7fffffffLL))
        ELL_FAIL(EINVAL, "dimension does not fit a 32-bit idx_t");
    if (rowsize > 0 && num_columns == 0 && row_end > row_begin)
        ELL_FAIL(EINVAL, "rowsize > 0 with zero columns");
    int err = check_device(&device);
    if (err) return err;
    ellspmv_cuda_matrix *A = new (std::nothrow) ellspmv_cuda_matrix();
    if (!A) ELL_FAIL(ENOMEM, "out of host memory");
    A->device = device;
    A->host_idx_bits = idx_width_bits;
    A->num_columns = num_columns;
    A->row_begin = row_begin;
    A->global_rows = global_rows;
    A->lay.num_rows = row_end - row_begin;
    A->lay.rowsize = (int)rowsize;
    *out = A;
    return 0;
}

}  // namespace

namespace ellspmv {
// one y <- beta*y + A*x on a single-GPU CSR handle: through the ELL view when there is one
int csr_launch(csrspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta, cudaStream_t stream)
{
    if (A->ell) return launch(A->ell, y_dev, x_dev, beta, nullptr, stream, 0, -1, nullptr);
    if (A->sell) {
        ELL_CK(sell_spmv(A->sell, A->fma, A->rowptr, A->cols, A->vals, A->idx_bits, x_dev, y_dev, A->d_ad, A->row_begin,
                         beta, stream));
        return 0;
    }
    CsrSpmvArgs args = {A->rowptr, A->cols, A->vals, x_dev, y_dev, A->num_rows, beta, A->d_ad, A->row_begin};
    ELL_CK(launch_csr_spmv(A->idx_bits, A->fma, A->kernel, args, stream));
    return 0;
}
int launch_shard(ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta,
                 const PushTargets *push, cudaStream_t stream)
{
    return launch(A, y_dev, x_dev, beta, push, stream, 0, -1);
}
int ensure_event_count(std::vector<cudaEvent_t> &ev, size_t n) { return ensure_events(ev, n); }

// The boundary slices of a shard under a push plan -- the ones that push to a peer or read a column
// outside the shard's own (16-entry-aligned) row range -- as runs of slice indices; recomputed only
// when the plan changes.  Everything else is interior: it neither needs a peer's data nor produces
// any, which is what both forms of the step hand-shake build on.
static int exchange_plan(ellspmv_cuda_matrix *A, const PushTargets *push, cudaStream_t stream)
{
    if (!A->d_remote) {
        const int64_t lo = (A->row_begin + 15) & ~(int64_t)15;
        const int64_t hi = (A->row_begin + A->lay.num_rows) & ~(int64_t)15;
        ELL_CK(cudaMalloc(&A->d_remote, (size_t)A->lay.num_slices));
        ELL_CK(cudaMalloc(&A->d_done, sizeof(unsigned)));
        ELL_CK(cudaMemsetAsync(A->d_done, 0, sizeof(unsigned), A->stream));
        ELL_CK(mark_remote_slices(A->dev_idx_bits, A->cols, A->lay, lo, hi, A->d_remote, A->stream));
        A->h_remote.resize((size_t)A->lay.num_slices);
        ELL_CK(cudaMemcpyAsync(A->h_remote.data(), A->d_remote, (size_t)A->lay.num_slices, cudaMemcpyDeviceToHost, A->stream));
        ELL_CK(cudaStreamSynchronize(A->stream));
        A->device_bytes += A->lay.num_slices + 4;
        A->sync_plan_valid = false;
    }
    bool same = A->sync_plan_valid && push && A->sync_plan.num_peers == push->num_peers;
    for (int p = 0; same && p < push->num_peers; p++)
        same = A->sync_plan.row_lo[p] == push->row_lo[p] && A->sync_plan.row_hi[p] == push->row_hi[p];
    if (same) return 0;
    const int64_t S = A->lay.slice_rows, per_warp = 32 * (int64_t)A->cfg.rows_per_thread;
    std::vector<unsigned char> flag((size_t)A->lay.num_slices);
    unsigned total = 0;
    int runs = 0;
    int64_t count = 0;
    long long lo[4] = {0, 0, 0, 0}, hi[4] = {0, 0, 0, 0};
    bool prev = false;
    for (int64_t sl = 0; sl < A->lay.num_slices; sl++) {
        bool b = A->h_remote[(size_t)sl] != 0;
        const int64_t g_lo = A->row_begin + sl * S, g_hi = g_lo + S;
        for (int p = 0; push && !b && p < push->num_peers; p++) b = g_lo < push->row_hi[p] && g_hi > push->row_lo[p];
        flag[(size_t)sl] = b ? 1 : 0;
        if (b) {
            const int64_t rows_here = (A->lay.num_rows - sl * S < S) ? A->lay.num_rows - sl * S : S;
            total += (unsigned)((rows_here + per_warp - 1) / per_warp);
            count++;
            if (!prev) { if (runs < 4) lo[runs] = sl; runs++; }
            if (runs <= 4) hi[runs - 1] = sl + 1;
        }
        prev = b;
    }
    A->sync_total = total;
    A->sync_boundary_slices = count;
    if (runs <= 4) {
        A->sync_num_ranges = runs;
        for (int i = 0; i < 4; i++) { A->sync_range_lo[i] = lo[i]; A->sync_range_hi[i] = hi[i]; }
    } else {
        A->sync_num_ranges = -1;
        if (!A->d_boundary) ELL_CK(cudaMalloc(&A->d_boundary, (size_t)A->lay.num_slices));
        ELL_CK(cudaMemcpyAsync(A->d_boundary, flag.data(), (size_t)A->lay.num_slices, cudaMemcpyHostToDevice, stream));
        ELL_CK(cudaStreamSynchronize(stream));          // `flag` goes out of scope
    }
    if (push) A->sync_plan = *push; else A->sync_plan.num_peers = 0;
    A->sync_plan_valid = push != nullptr;
    return 0;
}

int launch_shard_exchange(ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int beta,
                          const PushTargets *push, StepSync sync, cudaStream_t stream)
{
    const bool sliceable = A->cfg.kernel == ELLSPMV_CUDA_KERNEL_THREAD && !(A->cfg.variant & 1) && !A->sg && !A->cb &&
                           !A->sell && A->lay.rowsize > 0 && A->lay.num_rows > 0;
    int err;
    if (!sliceable) {
        // kernels that only run whole (long-row, staged gather, SELL ...): push, then the one-warp hand-shake
        if ((err = launch(A, y_dev, x_dev, beta, push, stream, 0, -1))) return err;
        ELL_CK(launch_peer_sync(sync, stream));
        return 0;
    }
    if ((err = exchange_plan(A, push, stream))) return err;
    if (fused_sync_capable(A)) {
        // opt-in: the hand-shake inside the SpMV kernel
        static const int poll_env = getenv("ELLSPMV_CUDA_SYNC_POLL_NS") ? atoi(getenv("ELLSPMV_CUDA_SYNC_POLL_NS")) : 100;
        static const int nowait_env = getenv("ELLSPMV_CUDA_SYNC_NOWAIT") ? atoi(getenv("ELLSPMV_CUDA_SYNC_NOWAIT")) : 0;
        sync.done = A->d_done;
        sync.poll_ns = (unsigned)(poll_env > 0 ? poll_env : 100);
        sync.debug_nowait = nowait_env;
        sync.num_ranges = A->sync_num_ranges;
        for (int i = 0; i < 4; i++) { sync.range_lo[i] = A->sync_range_lo[i]; sync.range_hi[i] = A->sync_range_hi[i]; }
        sync.table = A->d_boundary;
        sync.total_warps = A->sync_total;
        return launch(A, y_dev, x_dev, beta, push, stream, 0, -1, &sync);
    }
    // ELLSPMV_CUDA_SPLIT_EXCHANGE: 0 = one launch + hand-shake behind it, 1 = boundary, then interior
    // on one stream, 2 (default) = boundary and interior side by side on two streams
    static const int split_env = getenv("ELLSPMV_CUDA_NO_SPLIT_EXCHANGE") && atoi(getenv("ELLSPMV_CUDA_NO_SPLIT_EXCHANGE")) ? 0
                                 : getenv("ELLSPMV_CUDA_SPLIT_EXCHANGE") ? atoi(getenv("ELLSPMV_CUDA_SPLIT_EXCHANGE")) : 2;
    const int runs = A->sync_num_ranges;
    if (split_env <= 0 || runs < 1 || runs > 4 || A->sync_boundary_slices * 4 > A->lay.num_slices) {
        // no interior worth the name (a scattered matrix: every slice reads remote columns)
        if ((err = launch(A, y_dev, x_dev, beta, push, stream, 0, -1))) return err;
        ELL_CK(launch_peer_sync(sync, stream));
        return 0;
    }
    // Default: the step in two parts.  The boundary slices -- the only ones that need the peers'
    // pushes of the previous step and the only ones that push -- and their hand-shake with the
    // neighbouring ranks run on a second stream; the interior slices (nearly all of the work, no
    // dependence on any peer) run on the caller's stream at the same time.  The second stream keeps
    // its own order from step to step (boundary e, hand-shake e, boundary e+1 ...), so neither the
    // flag round trip over NVLink nor the skew between ranks is on the interior's critical path.
    // On return `stream` is ordered after ALL rows of y (not after the hand-shake: the halo of y
    // is for the next exchange call, which is).
    if (!A->side) {
        ELL_CK(cudaStreamCreateWithFlags(&A->side, cudaStreamNonBlocking));
        ELL_CK(cudaEventCreateWithFlags(&A->ev_start, cudaEventDisableTiming));
        ELL_CK(cudaEventCreateWithFlags(&A->ev_boundary, cudaEventDisableTiming));
        ELL_CK(cudaEventCreateWithFlags(&A->ev_handshake, cudaEventDisableTiming));
    }
    const bool side_by_side = split_env >= 2;
    cudaStream_t bs = side_by_side ? A->side : stream;
    if (side_by_side) {
        ELL_CK(cudaEventRecord(A->ev_start, stream));           // x complete, previous step's y rows consumed
        ELL_CK(cudaStreamWaitEvent(A->side, A->ev_start, 0));
    } else if (A->handshake_pending) {
        ELL_CK(cudaStreamWaitEvent(stream, A->ev_handshake, 0));
    }
    for (int i = 0; i < runs; i++)
        if ((err = launch(A, y_dev, x_dev, beta, push, bs, A->sync_range_lo[i], A->sync_range_hi[i] - A->sync_range_lo[i])))
            return err;
    ELL_CK(cudaEventRecord(A->ev_boundary, bs));
    if (!side_by_side) ELL_CK(cudaStreamWaitEvent(A->side, A->ev_boundary, 0));
    ELL_CK(launch_peer_sync(sync, A->side));
    ELL_CK(cudaEventRecord(A->ev_handshake, A->side));
    A->handshake_pending = true;
    int64_t at = 0;
    for (int i = 0; i <= runs; i++) {
        const int64_t end = i < runs ? A->sync_range_lo[i] : A->lay.num_slices;
        if (end > at && (err = launch(A, y_dev, x_dev, beta, nullptr, stream, at, end - at))) return err;
        if (i < runs) at = A->sync_range_hi[i];
    }
    if (side_by_side) ELL_CK(cudaStreamWaitEvent(stream, A->ev_boundary, 0));
    return 0;
}
void shard_x_range(const ellspmv_cuda_matrix *A, int64_t *lo, int64_t *hi) { x_range(A, lo, hi); }
// same for a CSR handle: the stored column range, plus its own global rows with csrgemvsd's diagonal
void csr_x_range(const csrspmv_cuda_matrix *A, int64_t *lo, int64_t *hi)
{
    int64_t a = A->min_col, b = A->max_col + 1;
    if (b <= a) { a = 0; b = 0; }
    if (A->d_ad && A->num_rows > 0) {
        const int64_t r0 = A->row_begin, r1 = A->row_begin + A->num_rows;
        if (b <= a) { a = r0; b = r1; }
        else { a = r0 < a ? r0 : a; b = r1 > b ? r1 : b; }
    }
    if (b > A->num_columns) b = A->num_columns;
    *lo = a; *hi = b;
}
}  // namespace ellspmv

extern "C" {

int ellspmv_cuda_version(void) { return ELLSPMV_CUDA_VERSION; }

const char *ellspmv_cuda_last_error(void) { return g_last_error; }

const char *ellspmv_cuda_strerror(int err)
{
    return err == 0 ? "success" : strerror(err);
}

int ellspmv_cuda_device_count(int *count)
{
    if (!count) ELL_FAIL(EINVAL, "count is NULL");
    *count = 0;
    cudaError_t e = cudaGetDeviceCount(count);
    if (e != cudaSuccess) { cudaGetLastError(); *count = 0; ELL_FAIL(ENODEV, "%s", cudaGetErrorString(e)); }
    return 0;
}

void ellspmv_cuda_free(ellspmv_cuda_matrix *A)
{
    if (!A) return;
    if (!A->shards.empty()) { group_free(A); return; }
    DeviceGuard g(A->device);
    if (A->stream) cudaStreamSynchronize(A->stream);
    for (cudaEvent_t e : A->events) cudaEventDestroy(e);
    if (A->vals) cudaFree(A->vals);
    if (A->cols) cudaFree(A->cols);
    if (A->d_minmax) cudaFree(A->d_minmax);
    if (A->d_ad) cudaFree(A->d_ad);
    if (A->d_rowlen) cudaFree(A->d_rowlen);
    if (A->d_remote) cudaFree(A->d_remote);
    if (A->d_boundary) cudaFree(A->d_boundary);
    if (A->side) { cudaStreamSynchronize(A->side); cudaStreamDestroy(A->side); }
    if (A->ev_start) cudaEventDestroy(A->ev_start);
    if (A->ev_boundary) cudaEventDestroy(A->ev_boundary);
    if (A->ev_handshake) cudaEventDestroy(A->ev_handshake);
    if (A->d_done) cudaFree(A->d_done);
    pattern_free(&A->pat);
    if (A->cb) cb_free(A->cb);
    if (A->sg) sg_free(A->sg);
    if (A->sell) sell_free(A->sell);
    if (A->d_x) cudaFree(A->d_x);
    if (A->d_y) cudaFree(A->d_y);
    if (A->stream) cudaStreamDestroy(A->stream);
    if (A->stream_out) cudaStreamDestroy(A->stream_out);
    if (A->stream_in) cudaStreamDestroy(A->stream_in);
    delete A;
}

int ellspmv_cuda_upload_shard(
    ellspmv_cuda_matrix **out, int idx_width_bits,
    int64_t global_rows, int64_t num_columns, int64_t rowsize,
    int64_t row_begin, int64_t row_end,
    const void *colidx, const double *a, int device, unsigned flags)
{
    int err = new_handle(out, idx_width_bits, global_rows, num_columns, rowsize, row_begin, row_end,
                         device, flags);
    if (err) return err;
    ellspmv_cuda_matrix *A = *out;
    DeviceGuard g(A->device);
    auto fail = [&](int e) { ellspmv_cuda_free(A); *out = nullptr; return e; };
    if ((err = configure(A, flags))) return fail(err);
    if ((err = alloc_matrix(A))) return fail(err);
    const int64_t rows = A->lay.num_rows, K = A->lay.rowsize;
    if (rows > 0 && K > 0) {
        if (!colidx || !a) { set_last_error("colidx or a is NULL"); return fail(EINVAL); }
        // stream the row-major arrays through two staging buffers
        const int ib = idx_width_bits / 8;
        int64_t chunk_rows = (64LL << 20) / (K * (8 + ib));
        if (chunk_rows < 1) chunk_rows = 1;
        if (chunk_rows > rows) chunk_rows = rows;
        double *sv[2] = {nullptr, nullptr};
        void *sc[2] = {nullptr, nullptr};
        cudaEvent_t done[2] = {nullptr, nullptr};
        cudaError_t ce = cudaSuccess;
        for (int b = 0; b < 2 && ce == cudaSuccess; b++) {
            ce = cudaMalloc(&sv[b], (size_t)chunk_rows * K * 8);
            if (ce == cudaSuccess) ce = cudaMalloc(&sc[b], (size_t)chunk_rows * K * ib);
            if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&done[b], cudaEventDisableTiming);
        }
        int b = 0;
        for (int64_t r0 = 0; r0 < rows && ce == cudaSuccess; r0 += chunk_rows, b ^= 1) {
            const int64_t n = (rows - r0 < chunk_rows) ? rows - r0 : chunk_rows;
            ce = cudaEventSynchronize(done[b]);   // staging buffer b free again
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(sv[b], a + r0 * K, (size_t)n * K * 8, cudaMemcpyDefault, A->stream);
            if (ce == cudaSuccess) ce = cudaMemcpyAsync(sc[b], (const char *)colidx + r0 * K * ib, (size_t)n * K * ib, cudaMemcpyDefault, A->stream);
            if (ce == cudaSuccess) ce = relayout_chunk(idx_width_bits, A->dev_idx_bits, sc[b], sv[b], A->cols, A->vals, A->lay, r0, n, A->d_minmax, A->stream);
            if (ce == cudaSuccess) ce = cudaEventRecord(done[b], A->stream);
        }
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(A->stream);
        for (int i = 0; i < 2; i++) {
            if (sv[i]) cudaFree(sv[i]);
            if (sc[i]) cudaFree(sc[i]);
            if (done[i]) cudaEventDestroy(done[i]);
        }
        if (ce != cudaSuccess) {
            set_last_error("upload: %s", cudaGetErrorString(ce));
            return fail(cuda_to_errno(ce));
        }
        if ((err = finish_minmax(A))) return fail(err);
        if ((err = build_patterns(A))) return fail(err);
        if ((err = build_column_blocks(A))) return fail(err);
    } else {
        cudaError_t ce = cudaStreamSynchronize(A->stream);
        if (ce != cudaSuccess) { set_last_error("upload: %s", cudaGetErrorString(ce)); return fail(cuda_to_errno(ce)); }
    }
    return 0;
}

int ellspmv_cuda_upload(
    ellspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns, int64_t rowsize,
    const void *colidx, const double *a, int num_gpus, unsigned flags)
{
    if (num_gpus < 1) ELL_FAIL(EINVAL, "num_gpus must be >= 1");
    if (num_gpus > 1)
        return group_upload(out, idx_width_bits, num_rows, num_columns, rowsize, colidx, a, num_gpus, flags);
    return ellspmv_cuda_upload_shard(out, idx_width_bits, num_rows, num_columns, rowsize,
                                     0, num_rows, colidx, a, -1, flags);
}

int ellspmv_cuda_upload_coo(
    ellspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a, int device, unsigned flags)
{
    if (!out) ELL_FAIL(EINVAL, "out is NULL");
    *out = nullptr;
    if (idx_width_bits != 32 && idx_width_bits != 64) ELL_FAIL(EINVAL, "idx_width_bits must be 32 or 64");
    if (num_rows < 0 || num_columns < 0 || num_nonzeros < 0) ELL_FAIL(EINVAL, "negative dimension");
    if (num_nonzeros > 0 && (!rowidx || !colidx || !a)) ELL_FAIL(EINVAL, "NULL COO array");
    int err = check_device(&device);
    if (err) return err;
    DeviceGuard g(device);
    const size_t nz = (size_t)(num_nonzeros > 0 ? num_nonzeros : 1), ib = (size_t)idx_width_bits / 8;
    void *d_ri = nullptr, *d_ci = nullptr;
    double *d_a = nullptr;
    cudaStream_t s = nullptr;
    CooEllJob job;
    int bad = 0;
    ellspmv_cuda_matrix *A = nullptr;
    cudaError_t ce = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_ri, nz * ib);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_ci, nz * ib);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_a, nz * 8);
    if (ce == cudaSuccess && num_nonzeros > 0) ce = cudaMemcpyAsync(d_ri, rowidx, nz * ib, cudaMemcpyDefault, s);
    if (ce == cudaSuccess && num_nonzeros > 0) ce = cudaMemcpyAsync(d_ci, colidx, nz * ib, cudaMemcpyDefault, s);
    if (ce == cudaSuccess && num_nonzeros > 0) ce = cudaMemcpyAsync(d_a, a, nz * 8, cudaMemcpyDefault, s);
    job.idx_bits = idx_width_bits;
    job.nnz = num_nonzeros; job.num_rows = num_rows; job.num_columns = num_columns;
    job.d_rowidx = d_ri; job.d_colidx = d_ci; job.d_a = d_a;
    if (ce == cudaSuccess) ce = coo_to_ell_phase1(job, s, &bad);
    err = 0;
    if (ce != cudaSuccess) { set_last_error("upload_coo: %s", cudaGetErrorString(ce)); err = cuda_to_errno(ce); }
    else if (bad) { set_last_error("upload_coo: row index outside [1, %lld]", (long long)num_rows); err = EINVAL; }
    if (!err) err = new_handle(&A, idx_width_bits, num_rows, num_columns, job.rowsize, 0, num_rows, device, flags);
    if (!err) err = configure(A, flags);
    if (!err) err = alloc_matrix(A);
    if (!err) {
        ce = cudaStreamSynchronize(A->stream);     // memsets of alloc_matrix before the scatter on `s`
        if (ce == cudaSuccess)
            ce = coo_to_ell_phase2(job, A->dev_idx_bits, A->cols, A->vals, A->lay, A->d_minmax, s, &bad);
        if (ce != cudaSuccess) { set_last_error("upload_coo: %s", cudaGetErrorString(ce)); err = cuda_to_errno(ce); }
        else if (bad) { set_last_error("upload_coo: column index outside [1, %lld]", (long long)num_columns); err = EINVAL; }
    }
    if (!err && A->lay.num_rows > 0 && A->lay.rowsize > 0) err = finish_minmax(A);
    if (!err) err = build_patterns(A);
    if (!err) err = build_column_blocks(A);
    coo_to_ell_release(job);
    cudaFree(d_ri); cudaFree(d_ci); cudaFree(d_a);
    if (s) cudaStreamDestroy(s);
    if (err) { if (A) ellspmv_cuda_free(A); return err; }
    *out = A;
    return 0;
}

int ellspmv_cuda_generate(
    ellspmv_cuda_matrix **out, int kind, const int64_t dims[3],
    const double vals[2], uint64_t seed, int idx_width_bits,
    int64_t row_begin, int64_t row_end, int device, unsigned flags)
{
    if (!dims) ELL_FAIL(EINVAL, "dims is NULL");
    int64_t rows, cols, K;
    double v[2] = {0.0, 0.0};
    if (vals) { v[0] = vals[0]; v[1] = vals[1]; }
    switch (kind) {
    case ELLSPMV_CUDA_GEN_LAPLACE2D:
        if (dims[0] < 1 || dims[1] < 1) ELL_FAIL(EINVAL, "laplace2d needs nx, ny >= 1");
        rows = cols = dims[0] * dims[1]; K = 5; break;
    case ELLSPMV_CUDA_GEN_STENCIL27:
        if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) ELL_FAIL(EINVAL, "stencil27 needs nx, ny, nz >= 1");
        rows = cols = dims[0] * dims[1] * dims[2]; K = 27; break;
    case ELLSPMV_CUDA_GEN_RANDOM:
        if (dims[0] < 0 || dims[1] < 1 || dims[2] < 0) ELL_FAIL(EINVAL, "random needs rows >= 0, cols >= 1, K >= 0");
        rows = dims[0]; cols = dims[1]; K = dims[2]; break;
    default:
        ELL_FAIL(EINVAL, "unknown generator kind %d", kind);
    }
    if (row_end < 0) row_end = rows;
    int err = new_handle(out, idx_width_bits, rows, cols, K, row_begin, row_end, device, flags);
    if (err) return err;
    ellspmv_cuda_matrix *A = *out;
    DeviceGuard g(A->device);
    auto fail = [&](int e) { ellspmv_cuda_free(A); *out = nullptr; return e; };
    if ((err = configure(A, flags))) return fail(err);
    if ((err = alloc_matrix(A))) return fail(err);
    if (A->lay.num_rows > 0 && K > 0) {
        cudaError_t ce = generate_sliced(kind, dims, v, seed, A->dev_idx_bits, A->cols, A->vals,
                                         A->lay, row_begin, A->d_minmax, A->stream);
        if (ce != cudaSuccess) { set_last_error("generate: %s", cudaGetErrorString(ce)); return fail(cuda_to_errno(ce)); }
        if ((err = finish_minmax(A))) return fail(err);
        if ((err = build_patterns(A))) return fail(err);
        if ((err = build_column_blocks(A))) return fail(err);
    } else {
        cudaStreamSynchronize(A->stream);
    }
    return 0;
}

int ellspmv_cuda_generate_sharded(
    ellspmv_cuda_matrix **out, int kind, const int64_t dims[3],
    const double vals[2], uint64_t seed, int idx_width_bits, int num_gpus, unsigned flags)
{
    if (num_gpus < 1) ELL_FAIL(EINVAL, "num_gpus must be >= 1");
    if (num_gpus == 1) return ellspmv_cuda_generate(out, kind, dims, vals, seed, idx_width_bits, 0, -1, -1, flags);
    return group_generate(out, kind, dims, vals, seed, idx_width_bits, num_gpus, flags);
}

int ellspmv_cuda_download(const ellspmv_cuda_matrix *A, void *colidx, double *a)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) return group_download(A, colidx, a);
    const int64_t rows = A->lay.num_rows, K = A->lay.rowsize;
    if (rows == 0 || K == 0) return 0;
    if (!colidx || !a) ELL_FAIL(EINVAL, "colidx or a is NULL");
    DeviceGuard g(A->device);
    const int ib = A->host_idx_bits / 8;
    int64_t chunk_rows = (64LL << 20) / (K * (8 + ib));
    if (chunk_rows < 1) chunk_rows = 1;
    if (chunk_rows > rows) chunk_rows = rows;
    double *sv = nullptr; void *sc = nullptr;
    ELL_CK(cudaMalloc(&sv, (size_t)chunk_rows * K * 8));
    cudaError_t ce = cudaMalloc(&sc, (size_t)chunk_rows * K * ib);
    for (int64_t r0 = 0; r0 < rows && ce == cudaSuccess; r0 += chunk_rows) {
        const int64_t n = (rows - r0 < chunk_rows) ? rows - r0 : chunk_rows;
        ce = unlayout_chunk(A->dev_idx_bits, A->host_idx_bits, A->cols, A->vals, sc, sv, A->lay, r0, n, A->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync(a + r0 * K, sv, (size_t)n * K * 8, cudaMemcpyDefault, A->stream);
        if (ce == cudaSuccess) ce = cudaMemcpyAsync((char *)colidx + r0 * K * ib, sc, (size_t)n * K * ib, cudaMemcpyDefault, A->stream);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(A->stream);
    }
    cudaFree(sv);
    if (sc) cudaFree(sc);
    if (ce != cudaSuccess) ELL_FAIL(cuda_to_errno(ce), "download: %s", cudaGetErrorString(ce));
    return 0;
}

int ellspmv_cuda_set_diagonal(ellspmv_cuda_matrix *A, const double *ad, int order)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) {
        // each shard takes its own rows of the diagonal
        for (ellspmv_cuda_matrix *S : A->shards) {
            int err = ellspmv_cuda_set_diagonal(S, ad ? ad + S->row_begin : nullptr, order);
            if (err) return err;
        }
        return 0;
    }
    if (order != 0 && order != 1) ELL_FAIL(EINVAL, "order must be 0 (ellgemvsd) or 1 (ellgemv16sd)");
    DeviceGuard g(A->device);
    if (!ad) {
        if (A->d_ad) { cudaFree(A->d_ad); A->d_ad = nullptr; }
        return 0;
    }
    // x[i] is read at the row's own (global) index: every shard row needs a column
    if (A->row_begin + A->lay.num_rows > A->num_columns)
        ELL_FAIL(EINVAL, "separate diagonal needs rows <= columns (ellgemvsd reads x[i] for every row i)");
    // the padded tail of the last slice is read with vector loads: allocate whole slices
    const size_t n = (size_t)(A->lay.padded_rows() > 0 ? A->lay.padded_rows() : 1);
    if (!A->d_ad) {
        ELL_CK(cudaMalloc(&A->d_ad, n * 8));
        A->device_bytes += (int64_t)n * 8;
    }
    ELL_CK(cudaMemsetAsync(A->d_ad, 0, n * 8, A->stream));
    if (A->lay.num_rows > 0)
        ELL_CK(cudaMemcpyAsync(A->d_ad, ad, (size_t)A->lay.num_rows * 8, cudaMemcpyDefault, A->stream));
    ELL_CK(cudaStreamSynchronize(A->stream));
    A->sd_order = order;
    return 0;
}

int ellspmv_cuda_get_info(const ellspmv_cuda_matrix *A, ellspmv_cuda_info *info)
{
    if (!A || !info) ELL_FAIL(EINVAL, "NULL argument");
    if (!A->shards.empty()) return group_info(A, info);
    memset(info, 0, sizeof(*info));
    info->num_rows = A->lay.num_rows;
    info->num_columns = A->num_columns;
    info->rowsize = A->lay.rowsize;
    info->row_begin = A->row_begin;
    info->global_rows = A->global_rows;
    info->idx_width_bits = A->host_idx_bits;
    info->dev_idx_bits = A->dev_idx_bits;
    info->slice_rows = A->lay.slice_rows;
    info->rows_per_thread = A->cfg.rows_per_thread;
    info->kernel = A->cfg.kernel;
    info->fma = A->cfg.fma ? 1 : 0;
    info->device = A->device;
    info->device_bytes = A->device_bytes;
    info->min_col = A->min_col;
    info->max_col = A->max_col;
    info->launches = A->launches;
    info->num_gpus = 1;
    info->pattern_rows = A->pat.covered * A->pat.group_rows - A->pat.explicit_lanes * A->cfg.rows_per_thread;
    info->exception_entries = A->pat.explicit_lanes * A->cfg.rows_per_thread * A->lay.rowsize;
    info->staged = A->sg ? A->staged_mode : 0;
    info->launches_per_spmv = A->sg ? sg_launches(A->sg) : (A->cb ? cb_blocks(A->cb) : (A->sell ? sell_launches(A->sell) : 1));
    info->sell_slots = A->sell ? sell_entries(A->sell) : 0;
    info->value_pattern_rows = A->pat.vpat ? info->pattern_rows : 0;
    info->pattern_id_bytes = A->pat.patlane ? A->pat.groups * 32 : (A->pat.max_explicit ? A->pat.groups * 8 : (A->pat.patid ? A->pat.groups : 0));
    info->tune_ms[0] = A->tune_ms[0];
    info->tune_ms[1] = A->tune_ms[1];
    info->long_rows = A->cfg.kernel == kKernelLongRow ? A->lay.rowsize : 0;
    return 0;
}

int ellspmv_cuda_spmv_device(
    ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int mode, void *stream)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) ELL_FAIL(EINVAL, "spmv_device needs a single-GPU handle (this one spans %d GPUs)", (int)A->shards.size());
    if (mode != ELLSPMV_CUDA_ACCUMULATE && mode != ELLSPMV_CUDA_OVERWRITE)
        ELL_FAIL(EINVAL, "spmv_device: mode must be ACCUMULATE or OVERWRITE");
    if (A->lay.num_rows > 0 && (!y_dev || (!x_dev && A->lay.rowsize > 0)))
        ELL_FAIL(EINVAL, "NULL device vector");
    DeviceGuard g(A->device);
    return launch(A, y_dev, x_dev, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, nullptr, (cudaStream_t)stream);
}

int ellspmv_cuda_spmv_push(
    ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int mode,
    int num_peers, double *const *peer_x,
    const int64_t *peer_row_lo, const int64_t *peer_row_hi, void *stream)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) ELL_FAIL(EINVAL, "spmv_push needs a single-GPU handle");
    if (mode != ELLSPMV_CUDA_ACCUMULATE && mode != ELLSPMV_CUDA_OVERWRITE)
        ELL_FAIL(EINVAL, "spmv_push: mode must be ACCUMULATE or OVERWRITE");
    if (num_peers < 0 || num_peers > kMaxPeers) ELL_FAIL(EINVAL, "num_peers must be 0..%d", kMaxPeers);
    if (num_peers > 0 && (!peer_x || !peer_row_lo || !peer_row_hi)) ELL_FAIL(EINVAL, "NULL peer arrays");
    if (A->lay.num_rows > 0 && (!y_dev || (!x_dev && A->lay.rowsize > 0)))
        ELL_FAIL(EINVAL, "NULL device vector");
    PushTargets pt = {};
    pt.num_peers = num_peers;
    for (int p = 0; p < num_peers; p++) {
        if (!peer_x[p]) ELL_FAIL(EINVAL, "peer_x[%d] is NULL", p);
        pt.x[p] = peer_x[p];
        pt.row_lo[p] = peer_row_lo[p];
        pt.row_hi[p] = peer_row_hi[p];
    }
    DeviceGuard g(A->device);
    return launch(A, y_dev, x_dev, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, &pt, (cudaStream_t)stream);
}

int ellspmv_cuda_spmv_exchange(
    ellspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int mode,
    int num_peers, double *const *peer_x, const int64_t *peer_row_lo, const int64_t *peer_row_hi,
    int rank, int num_sync, const int *sync_ranks, int64_t *const *sync_flags,
    int64_t *local_flags, int64_t epoch, void *stream)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) ELL_FAIL(EINVAL, "spmv_exchange needs a single-GPU handle");
    if (mode != ELLSPMV_CUDA_ACCUMULATE && mode != ELLSPMV_CUDA_OVERWRITE)
        ELL_FAIL(EINVAL, "spmv_exchange: mode must be ACCUMULATE or OVERWRITE");
    if (num_peers < 0 || num_peers > kMaxPeers) ELL_FAIL(EINVAL, "num_peers must be 0..%d", kMaxPeers);
    if (num_peers > 0 && (!peer_x || !peer_row_lo || !peer_row_hi)) ELL_FAIL(EINVAL, "NULL peer arrays");
    if (num_sync < 0 || num_sync > kMaxPeers) ELL_FAIL(EINVAL, "num_sync must be 0..%d", kMaxPeers);
    if (num_sync > 0 && (!sync_ranks || !sync_flags)) ELL_FAIL(EINVAL, "NULL sync arrays");
    if (!local_flags) ELL_FAIL(EINVAL, "local_flags is NULL");
    if (rank < 0 || rank >= kMaxRanks) ELL_FAIL(EINVAL, "rank %d out of range (max %d ranks)", rank, kMaxRanks);
    if (epoch < 1) ELL_FAIL(EINVAL, "epoch must be >= 1");
    if (A->lay.num_rows > 0 && (!y_dev || (!x_dev && A->lay.rowsize > 0)))
        ELL_FAIL(EINVAL, "NULL device vector");
    PushTargets pt = {};
    pt.num_peers = num_peers;
    for (int p = 0; p < num_peers; p++) {
        if (!peer_x[p]) ELL_FAIL(EINVAL, "peer_x[%d] is NULL", p);
        pt.x[p] = peer_x[p];
        pt.row_lo[p] = peer_row_lo[p];
        pt.row_hi[p] = peer_row_hi[p];
    }
    StepSync sy = {};
    sy.local_flags = reinterpret_cast<long long *>(local_flags);
    sy.num_peers = num_sync;
    sy.rank = rank;
    sy.epoch = (long long)epoch;
    sy.error = reinterpret_cast<int *>(local_flags + kMaxRanks);
    for (int p = 0; p < num_sync; p++) {
        if (!sync_flags[p]) ELL_FAIL(EINVAL, "sync_flags[%d] is NULL", p);
        if (sync_ranks[p] < 0 || sync_ranks[p] >= kMaxRanks) ELL_FAIL(EINVAL, "sync_ranks[%d] out of range", p);
        sy.peer_flags[p] = reinterpret_cast<long long *>(sync_flags[p]);
        sy.peer_rank[p] = sync_ranks[p];
    }
    DeviceGuard g(A->device);
    return launch_shard_exchange(A, y_dev, x_dev, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, &pt, sy,
                                 (cudaStream_t)stream);
}

int ellspmv_cuda_spmv(
    ellspmv_cuda_matrix *A, double *y, const double *x,
    int repeat, int mode, double *seconds)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (repeat < 0) ELL_FAIL(EINVAL, "repeat < 0");
    if (mode != ELLSPMV_CUDA_ACCUMULATE && mode != ELLSPMV_CUDA_OVERWRITE && mode != ELLSPMV_CUDA_ITERATE)
        ELL_FAIL(EINVAL, "unknown mode %d", mode);
    if (!A->shards.empty()) return group_spmv(A, y, x, repeat, mode, seconds);
    const int64_t rows = A->lay.num_rows, ncols = A->num_columns;
    if ((rows > 0 && !y) || (ncols > 0 && !x)) ELL_FAIL(EINVAL, "NULL host vector");
    if (mode == ELLSPMV_CUDA_ITERATE && !(A->row_begin == 0 && rows == A->global_rows && rows == ncols))
        ELL_FAIL(EINVAL, "ITERATE needs a square, unsharded matrix");
    if (repeat == 0) return 0;
    DeviceGuard g(A->device);
    int err = ensure_vectors(A);
    if (err) return err;
    if (repeat == 1 && mode != ELLSPMV_CUDA_ITERATE && pipelined_host_call_applies(A))
        return spmv_pipelined(A, y, x, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, seconds);
    if ((err = ensure_events(A->events, (size_t)repeat + 1))) return err;
    cudaStream_t s = A->stream;
    int64_t xlo = 0, xhi = ncols;
    if (mode != ELLSPMV_CUDA_ITERATE) x_range(A, &xlo, &xhi);      // ITERATE: every entry becomes an output
    if (xhi > xlo) ELL_CK(cudaMemcpyAsync(A->d_x + xlo, x + xlo, (size_t)(xhi - xlo) * 8, cudaMemcpyDefault, s));
    if (mode == ELLSPMV_CUDA_ACCUMULATE && rows > 0)
        ELL_CK(cudaMemcpyAsync(A->d_y, y, (size_t)rows * 8, cudaMemcpyDefault, s));
    double *cur = A->d_x, *nxt = A->d_y;
    ELL_CK(cudaEventRecord(A->events[0], s));
    for (int r = 0; r < repeat; r++) {
        if (mode == ELLSPMV_CUDA_ITERATE) {
            if ((err = launch(A, nxt, cur, 0, nullptr, s))) return err;
            double *t = cur; cur = nxt; nxt = t;
        } else {
            const int beta = mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0;
            if ((err = launch(A, A->d_y, A->d_x, beta, nullptr, s))) return err;
        }
        ELL_CK(cudaEventRecord(A->events[(size_t)r + 1], s));
    }
    const double *result = (mode == ELLSPMV_CUDA_ITERATE) ? cur : A->d_y;
    if (rows > 0) ELL_CK(cudaMemcpyAsync(y, result, (size_t)rows * 8, cudaMemcpyDefault, s));
    ELL_CK(cudaStreamSynchronize(s));
    if (seconds) {
        for (int r = 0; r < repeat; r++) {
            float ms = 0.f;
            ELL_CK(cudaEventElapsedTime(&ms, A->events[(size_t)r], A->events[(size_t)r + 1]));
            seconds[r] = (double)ms * 1e-3;
        }
    }
    return 0;
}

/* ---- CSR ---------------------------------------------------------------- */

void csrspmv_cuda_free(csrspmv_cuda_matrix *A)
{
    if (!A) return;
    if (!A->shards.empty()) { csr_group_free(A); return; }
    DeviceGuard g(A->device);
    if (A->stream) cudaStreamSynchronize(A->stream);
    for (cudaEvent_t e : A->events) cudaEventDestroy(e);
    if (A->rowptr) cudaFree(A->rowptr);
    if (A->d_ad) cudaFree(A->d_ad);
    if (A->d_scratch) cudaFree(A->d_scratch);
    if (A->ell) ellspmv_cuda_free(A->ell);
    if (A->sell) sell_free(A->sell);
    if (A->cols) cudaFree(A->cols);
    if (A->vals) cudaFree(A->vals);
    if (A->d_x) cudaFree(A->d_x);
    if (A->d_y) cudaFree(A->d_y);
    if (A->stream) cudaStreamDestroy(A->stream);
    delete A;
}

static int csr_new(csrspmv_cuda_matrix **out, int idx_width_bits, int64_t num_rows,
                   int64_t num_columns, int64_t csrsize, int device, unsigned flags)
{
    if (!out) ELL_FAIL(EINVAL, "out is NULL");
    *out = nullptr;
    if (idx_width_bits != 32 && idx_width_bits != 64) ELL_FAIL(EINVAL, "idx_width_bits must be 32 or 64");
    if (num_rows < 0 || num_columns < 0 || csrsize < 0) ELL_FAIL(EINVAL, "negative dimension");
    int err = check_device(&device);
    if (err) return err;
    csrspmv_cuda_matrix *A = new (std::nothrow) csrspmv_cuda_matrix();
    if (!A) ELL_FAIL(ENOMEM, "out of host memory");
    A->device = device;
    A->idx_bits = idx_width_bits;
    A->num_rows = num_rows;
    A->num_columns = num_columns;
    A->csrsize = csrsize;
    A->flags = flags;
    A->fma = (flags & ELLSPMV_CUDA_FMA) != 0;
    int kernel = flags & ELLSPMV_CUDA_KERNEL_MASK;
    // 0 = auto: bit-exact, scalar for balanced rows / stream for ragged ones (csr_pick_kernel);
    // 1 = stream, 2 = vector (tolerance), 3 = scalar
    A->auto_kernel = kernel == ELLSPMV_CUDA_KERNEL_AUTO;
    A->kernel = kernel == ELLSPMV_CUDA_KERNEL_WARP ? ELLSPMV_CUDA_KERNEL_WARP
              : (kernel == 3 ? 3 : (kernel == 5 ? 5 : ELLSPMV_CUDA_KERNEL_THREAD));      // 5: SELL-128-sigma (sell.cu)
    *out = A;
    DeviceGuard g(device);
    auto fail = [&](cudaError_t ce) {
        set_last_error("csr alloc: %s", cudaGetErrorString(ce));
        csrspmv_cuda_free(A); *out = nullptr; return cuda_to_errno(ce);
    };
    cudaError_t ce;
    const size_t nz = (size_t)(csrsize > 0 ? csrsize : 1);
    if ((ce = cudaMalloc(&A->rowptr, (size_t)(num_rows + 1) * 8)) != cudaSuccess) return fail(ce);
    if ((ce = cudaMalloc(&A->cols, nz * (idx_width_bits / 8))) != cudaSuccess) return fail(ce);
    if ((ce = cudaMalloc(&A->vals, nz * 8)) != cudaSuccess) return fail(ce);
    if ((ce = cudaMalloc(&A->d_x, (size_t)(num_columns > 0 ? num_columns : 1) * 8)) != cudaSuccess) return fail(ce);
    if ((ce = cudaMalloc(&A->d_y, (size_t)(num_rows > 0 ? num_rows : 1) * 8)) != cudaSuccess) return fail(ce);
    if ((ce = cudaMalloc(&A->d_scratch, 40)) != cudaSuccess) return fail(ce);
    if ((ce = cudaStreamCreateWithFlags(&A->stream, cudaStreamNonBlocking)) != cudaSuccess) return fail(ce);
    A->device_bytes = (num_rows + 1) * 8 + (int64_t)nz * (8 + idx_width_bits / 8) + (num_columns + num_rows) * 8;
    return 0;
}

// KERNEL_AUTO on a CSR matrix with balanced rows: keep a sliced-ELL view of the same entries
// (width = the longest row) and run the launches through the ELL kernels -- coalesced streams,
// offset patterns, L2 prefetch and the staged gather for scattered columns all apply, where the
// native CSR kernels below gather with one thread per row.  The view carries the rows' lengths:
// a slot past a row's end is never touched arithmetically, so the result is csrgemv's
// (csrspmv.c:1588-1593) bit for bit, non-finite x included.  Rows of one length need no
// length array at all (the random matrix of BASELINE config 4, every row K entries).
// Skipped when the padding would exceed 25 % of the entries, a row is longer than 1024 (unless
// the matrix has at most 32768 rows: the long-row kernel then runs the view), or
// memory is short; the CSR arrays stay on the device either way (download, fallback).
static int csr_build_ell_view(csrspmv_cuda_matrix *A)
{
    if (A->num_rows <= 0 || A->csrsize <= 0 || A->max_row_len <= 0) return 0;
    if (A->max_row_len > (A->num_rows <= 32768 ? (1 << 24) : 1024)) return 0;    // long rows only with few of them (long-row kernel)
    if (getenv("CSRSPMV_CUDA_NO_ELL_VIEW")) return 0;
    const int64_t K = A->max_row_len;
    const int64_t padded = A->num_rows * K;
    if (padded > A->csrsize + A->csrsize / 4 + 4096) return 0;
    const bool uniform = A->min_row_len == A->max_row_len;
    size_t free_b = 0, total_b = 0;
    ELL_CK(cudaMemGetInfo(&free_b, &total_b));
    if ((int64_t)free_b < padded * (8 + A->idx_bits / 8) + A->num_rows * 4 + (1LL << 30)) return 0;
    unsigned flags = A->flags & (ELLSPMV_CUDA_FMA | ELLSPMV_CUDA_WIDE_INDEX | ELLSPMV_CUDA_NO_PATTERN | ELLSPMV_CUDA_PATTERN_MASKS | ELLSPMV_CUDA_NO_PATTERN_LANES |
                                 ELLSPMV_CUDA_NO_STAGED_GATHER | ELLSPMV_CUDA_STAGED_GATHER | ELLSPMV_CUDA_L2_PERSIST_X);
    // kernel of the view: the same nnz-per-row switch as an ELL upload (configure); both kernels
    // honour the row lengths, the thread-per-row one in its R = 1 / explicit-index form
    if (K >= 64 && A->num_rows <= 32768 && !(flags & ELLSPMV_CUDA_STAGED_GATHER)) flags |= kKernelLongRow;
    else {
        flags |= ELLSPMV_CUDA_KERNEL_THREAD;
        // rows with their own lengths: one row per thread; offset patterns apply all the same (a
        // boundary row of a stencil is one more kind of row: the unused slots repeat its last
        // column, csr_to_sliced), only the lane-mask form is not instantiated for them
        if (!uniform) flags = (flags & ~(unsigned)ELLSPMV_CUDA_PATTERN_MASKS) | (1u << ELLSPMV_CUDA_ROWS_PER_THREAD_SHIFT);
    }
    ellspmv_cuda_matrix *V = nullptr;
    int err = new_handle(&V, A->idx_bits, A->row_begin + A->num_rows, A->num_columns, K, A->row_begin,
                         A->row_begin + A->num_rows, A->device, flags);
    if (err) return err;
    auto fail = [&](int e) { ellspmv_cuda_free(V); return e; };
    if ((err = configure(V, flags))) return fail(err);
    V->kernel_auto = true;                         // the view may still try the staged gather (build_column_blocks)
    if ((err = alloc_matrix(V))) {
        if (err == ENOMEM) { ellspmv_cuda_free(V); cudaGetLastError(); return 0; }
        return fail(err);
    }
    cudaError_t ce = cudaSuccess;
    if (!uniform) ce = cudaMalloc(&V->d_rowlen, (size_t)V->lay.padded_rows() * sizeof(int));
    if (ce == cudaSuccess && !uniform) ce = cudaMemsetAsync(V->d_rowlen, 0, (size_t)V->lay.padded_rows() * sizeof(int), V->stream);
    if (ce == cudaSuccess)
        ce = csr_to_sliced(A->idx_bits, V->dev_idx_bits, A->rowptr, A->cols, A->vals, V->cols, V->vals, V->d_rowlen,
                           V->lay, V->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(V->stream);
    if (ce != cudaSuccess) { set_last_error("csr -> ell view: %s", cudaGetErrorString(ce)); return fail(cuda_to_errno(ce)); }
    if (!uniform) V->device_bytes += V->lay.padded_rows() * (int64_t)sizeof(int);
    V->min_col = A->min_col;                       // checked by csr_inspect already
    V->max_col = A->max_col;
    warm_kernels(V);
    if ((err = build_patterns(V))) return fail(err);
    if ((err = build_column_blocks(V))) return fail(err);
    A->ell = V;
    A->device_bytes += V->device_bytes;
    return 0;
}

// KERNEL_AUTO on a CSR matrix whose rows are NOT balanced (the view above would be mostly padding):
// SELL-128-sigma (sell.cu) -- rows sorted by length inside windows of 4096, a width per slice,
// thread per row in entry order, rows beyond 4096 entries one CTA each.  Bit-exact like every
// default path; the entry-balancing of the reference's csrgemvnz (csrspmv.c:1698-1740) is done
// here by the layout instead of by atomics.  Falls back to the native kernels when memory is short.
static int csr_build_sell(csrspmv_cuda_matrix *A)
{
    if (A->num_rows <= 0 || A->csrsize <= 0) { if (A->kernel == 5) A->kernel = 3; return 0; }
    if (getenv("CSRSPMV_CUDA_NO_SELL") && A->auto_kernel) return 0;
    size_t free_b = 0, total_b = 0;
    ELL_CK(cudaMemGetInfo(&free_b, &total_b));
    const int dev_bits = (A->idx_bits == 64 && !(A->flags & ELLSPMV_CUDA_WIDE_INDEX) && A->num_columns < (1LL << 31)) ? 32 : A->idx_bits;
    const int64_t need = (A->csrsize + A->csrsize / 2) * (8 + dev_bits / 8) + A->num_rows * 16 + (1LL << 30);
    if ((int64_t)free_b < need) { if (A->kernel == 5) A->kernel = 3; return 0; }
    cudaError_t ce = sell_build_csr(&A->sell, A->idx_bits, dev_bits, A->num_rows, A->rowptr, A->cols, A->vals, A->stream);
    if (ce == cudaErrorMemoryAllocation) { cudaGetLastError(); A->sell = nullptr; if (A->kernel == 5) A->kernel = 3; return 0; }
    if (ce != cudaSuccess) { set_last_error("csr -> sell: %s", cudaGetErrorString(ce)); return cuda_to_errno(ce); }
    if (A->sell) {
        A->kernel = 5;
        A->device_bytes += sell_bytes(A->sell);
    }
    return 0;
}

// bit-exact CSR kernels: thread-per-row (scalar) wins when the rows are balanced
// (37 vs 48 ms on BASELINE config 4), the smem-staged stream kernel when they are ragged
static int csr_pick_kernel(csrspmv_cuda_matrix *A)
{
    // the same checks an ELL upload makes: a bad rowptr or column index is EINVAL here,
    // never an out-of-bounds gather in the kernels
    CsrInspection in;
    ELL_CK(csr_inspect(A->idx_bits, A->rowptr, A->cols, A->num_rows, A->csrsize, A->d_scratch, &in, A->stream));
    if (in.bad_rows > 0)
        ELL_FAIL(EINVAL, "rowptr is not non-decreasing (%lld row(s) end before they start)", (long long)in.bad_rows);
    if (A->csrsize > 0 && (in.min_col < 0 || in.max_col >= A->num_columns))
        ELL_FAIL(EINVAL, "column index out of range: [%lld, %lld] with %lld columns", (long long)in.min_col,
                 (long long)in.max_col, (long long)A->num_columns);
    A->max_row_len = in.max_row_len;
    A->min_row_len = in.min_row_len;
    A->min_col = in.min_col;
    A->max_col = in.max_col;
    if (!A->auto_kernel) return A->kernel == 5 ? csr_build_sell(A) : 0;
    const int64_t avg = A->num_rows > 0 ? (A->csrsize + A->num_rows - 1) / A->num_rows : 0;
    A->kernel = (A->max_row_len <= 4 * avg + 16) ? 3 : ELLSPMV_CUDA_KERNEL_THREAD;
    // ELLSPMV_CUDA_FMA: the stream kernel parks ROUNDED products and cannot contract; so that the
    // bits under FMA do not depend on the row-length distribution (or differ between the shards of
    // a group), AUTO then always takes a sequential-fma kernel: the ELL view or the scalar kernel
    if (A->fma) A->kernel = 3;
    int err = csr_build_ell_view(A);
    if (err || A->ell) return err;
    return csr_build_sell(A);
}

int csrspmv_cuda_upload(
    csrspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns,
    const int64_t *rowptr, const void *colidx, const double *a,
    int num_gpus, unsigned flags)
{
    if (num_gpus < 1) ELL_FAIL(EINVAL, "num_gpus must be >= 1");
    if (num_gpus > 1)
        return csr_group_upload(out, idx_width_bits, num_rows, num_columns, rowptr, colidx, a, num_gpus, flags);
    return csr_upload_on(out, idx_width_bits, num_rows, num_columns, rowptr, colidx, a, -1, flags);
}

int csrspmv_cuda_upload_coo(
    csrspmv_cuda_matrix **out, int idx_width_bits,
    int64_t num_rows, int64_t num_columns, int64_t num_nonzeros,
    const void *rowidx, const void *colidx, const double *a, unsigned flags)
{
    if (!out) ELL_FAIL(EINVAL, "out is NULL");
    *out = nullptr;
    if (num_nonzeros > 0 && (!rowidx || !colidx || !a)) ELL_FAIL(EINVAL, "NULL COO array");
    int err = csr_new(out, idx_width_bits, num_rows, num_columns, num_nonzeros, -1, flags);
    if (err) return err;
    csrspmv_cuda_matrix *A = *out;
    DeviceGuard g(A->device);
    const size_t nz = (size_t)(num_nonzeros > 0 ? num_nonzeros : 1), ib = (size_t)idx_width_bits / 8;
    void *d_ri = nullptr, *d_ci = nullptr;
    double *d_a = nullptr;
    int bad = 0;
    cudaStream_t s = A->stream;
    cudaError_t ce = cudaMalloc(&d_ri, nz * ib);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_ci, nz * ib);
    if (ce == cudaSuccess) ce = cudaMalloc(&d_a, nz * 8);
    if (ce == cudaSuccess && num_nonzeros > 0) ce = cudaMemcpyAsync(d_ri, rowidx, nz * ib, cudaMemcpyDefault, s);
    if (ce == cudaSuccess && num_nonzeros > 0) ce = cudaMemcpyAsync(d_ci, colidx, nz * ib, cudaMemcpyDefault, s);
    if (ce == cudaSuccess && num_nonzeros > 0) ce = cudaMemcpyAsync(d_a, a, nz * 8, cudaMemcpyDefault, s);
    if (ce == cudaSuccess)
        ce = coo_to_csr(idx_width_bits, d_ri, d_ci, d_a, num_nonzeros, num_rows, num_columns, A->rowptr, A->cols,
                        A->vals, s, &bad);
    cudaFree(d_ri); cudaFree(d_ci); cudaFree(d_a);
    if (ce != cudaSuccess || bad) {
        if (ce != cudaSuccess) set_last_error("csr upload_coo: %s", cudaGetErrorString(ce));
        else set_last_error("csr upload_coo: row or column index out of range");
        csrspmv_cuda_free(A); *out = nullptr;
        return ce != cudaSuccess ? cuda_to_errno(ce) : EINVAL;
    }
    if ((err = csr_pick_kernel(A))) { csrspmv_cuda_free(A); *out = nullptr; return err; }
    return 0;
}

int csrspmv_cuda_generate(
    csrspmv_cuda_matrix **out, int kind, const int64_t dims[3],
    const double vals[2], uint64_t seed, int idx_width_bits,
    int device, unsigned flags)
{
    if (!dims) ELL_FAIL(EINVAL, "dims is NULL");
    int64_t rows = 0, cols = 0, nnz = 0;
    if (kind == ELLSPMV_CUDA_GEN_RANDOM) {
        if (dims[0] < 0 || dims[1] < 1 || dims[2] < 0) ELL_FAIL(EINVAL, "random needs rows >= 0, cols >= 1, K >= 0");
        rows = dims[0]; cols = dims[1]; nnz = dims[0] * dims[2];
    } else if (kind == ELLSPMV_CUDA_GEN_LAPLACE2D || kind == ELLSPMV_CUDA_GEN_STENCIL27) {
        // the stencils as csr_from_coo stores them: no padding, boundary rows are shorter
        if (!vals) ELL_FAIL(EINVAL, "vals is NULL");
        const int nd = kind == ELLSPMV_CUDA_GEN_LAPLACE2D ? 2 : 3;
        rows = 1;
        for (int d = 0; d < nd; d++) {
            if (dims[d] < 1) ELL_FAIL(EINVAL, "grid dimensions must be >= 1");
            if (rows > (1LL << 40) / dims[d]) ELL_FAIL(EINVAL, "grid too large");
            rows *= dims[d];
        }
        cols = rows;
        nnz = csr_stencil_nnz(kind, dims);
    } else {
        ELL_FAIL(EINVAL, "unknown generator kind %d", kind);
    }
    if (idx_width_bits == 32 && cols > 0x7fffffffLL) ELL_FAIL(EINVAL, "32-bit indices cannot address %lld columns", (long long)cols);
    int err = csr_new(out, idx_width_bits, rows, cols, nnz, device, flags);
    if (err) return err;
    csrspmv_cuda_matrix *A = *out;
    DeviceGuard g(A->device);
    cudaError_t ce = kind == ELLSPMV_CUDA_GEN_RANDOM
        ? generate_csr_random(dims, seed, idx_width_bits, A->rowptr, A->cols, A->vals, A->stream)
        : generate_csr_stencil(kind, dims, vals, idx_width_bits, A->rowptr, A->cols, A->vals, A->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(A->stream);
    if (ce != cudaSuccess) {
        set_last_error("csr generate: %s", cudaGetErrorString(ce));
        csrspmv_cuda_free(A); *out = nullptr;
        return cuda_to_errno(ce);
    }
    if ((err = csr_pick_kernel(A))) { csrspmv_cuda_free(A); *out = nullptr; return err; }
    return 0;
}

int csrspmv_cuda_set_diagonal(csrspmv_cuda_matrix *A, const double *ad)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) {
        for (size_t p = 0; p < A->shards.size(); p++) {
            int err = csrspmv_cuda_set_diagonal(A->shards[p], ad ? ad + A->row_lo[p] : nullptr);
            if (err) return err;
        }
        return 0;
    }
    DeviceGuard g(A->device);
    if (!ad) {
        if (A->d_ad) { cudaFree(A->d_ad); A->d_ad = nullptr; }
        if (A->ell) return ellspmv_cuda_set_diagonal(A->ell, nullptr, 0);
        return 0;
    }
    // x[i] is read at the row's own GLOBAL index (a shard's rows start at row_begin)
    if (A->row_begin + A->num_rows > A->num_columns)
        ELL_FAIL(EINVAL, "separate diagonal needs rows <= columns (csrgemvsd reads x[i] for every row i)");
    const size_t n = (size_t)(A->num_rows > 0 ? A->num_rows : 1);
    if (!A->d_ad) {
        ELL_CK(cudaMalloc(&A->d_ad, n * 8));
        A->device_bytes += (int64_t)n * 8;
    }
    if (A->num_rows > 0)
        ELL_CK(cudaMemcpyAsync(A->d_ad, ad, (size_t)A->num_rows * 8, cudaMemcpyDefault, A->stream));
    ELL_CK(cudaStreamSynchronize(A->stream));
    // csrgemvsd (csrspmv.c:1622-1627) is ellgemvsd's order: the slots summed from 0, then ad*x + yi
    if (A->ell) return ellspmv_cuda_set_diagonal(A->ell, A->d_ad, 0);
    return 0;
}

int csrspmv_cuda_spmv_device(
    csrspmv_cuda_matrix *A, double *y_dev, const double *x_dev, int mode, void *stream)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) ELL_FAIL(EINVAL, "spmv_device needs a single-GPU handle");
    if (mode != ELLSPMV_CUDA_ACCUMULATE && mode != ELLSPMV_CUDA_OVERWRITE)
        ELL_FAIL(EINVAL, "mode must be ACCUMULATE or OVERWRITE");
    if (A->num_rows > 0 && (!y_dev || !x_dev)) ELL_FAIL(EINVAL, "NULL device vector");
    DeviceGuard g(A->device);
    return csr_launch(A, y_dev, x_dev, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, (cudaStream_t)stream);
}

int csrspmv_cuda_spmv(
    csrspmv_cuda_matrix *A, double *y, const double *x,
    int repeat, int mode, double *seconds)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (repeat < 0) ELL_FAIL(EINVAL, "repeat < 0");
    if (mode != ELLSPMV_CUDA_ACCUMULATE && mode != ELLSPMV_CUDA_OVERWRITE)
        ELL_FAIL(EINVAL, "mode must be ACCUMULATE or OVERWRITE");
    if ((A->num_rows > 0 && !y) || (A->num_columns > 0 && !x)) ELL_FAIL(EINVAL, "NULL host vector");
    if (repeat == 0) return 0;
    if (!A->shards.empty()) return csr_group_spmv(A, y, x, repeat, mode, seconds);
    DeviceGuard g(A->device);
    if (repeat == 1 && A->ell && pipelined_host_call_applies(A->ell)) {
        // one launch through the sliced-ELL view: the ELL host call's pipeline (upload, kernel and download
        // of row chunks on three streams) on this handle's vectors, lent to the view for the call
        ellspmv_cuda_matrix *V = A->ell;
        V->d_x = A->d_x; V->d_y = A->d_y;
        V->vec_len = A->num_rows > A->num_columns ? A->num_rows : A->num_columns;
        const int e = spmv_pipelined(V, y, x, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, seconds);
        V->d_x = V->d_y = nullptr;
        V->vec_len = 0;
        return e;
    }
    int err = ensure_events(A->events, (size_t)repeat + 1);
    if (err) return err;
    cudaStream_t s = A->stream;
    int64_t xlo, xhi;
    csr_x_range(A, &xlo, &xhi);
    if (xhi > xlo) ELL_CK(cudaMemcpyAsync(A->d_x + xlo, x + xlo, (size_t)(xhi - xlo) * 8, cudaMemcpyDefault, s));
    if (mode == ELLSPMV_CUDA_ACCUMULATE && A->num_rows > 0)
        ELL_CK(cudaMemcpyAsync(A->d_y, y, (size_t)A->num_rows * 8, cudaMemcpyDefault, s));
    ELL_CK(cudaEventRecord(A->events[0], s));
    for (int r = 0; r < repeat; r++) {
        if ((err = csr_launch(A, A->d_y, A->d_x, mode == ELLSPMV_CUDA_ACCUMULATE ? 1 : 0, s))) return err;
        ELL_CK(cudaEventRecord(A->events[(size_t)r + 1], s));
    }
    if (A->num_rows > 0) ELL_CK(cudaMemcpyAsync(y, A->d_y, (size_t)A->num_rows * 8, cudaMemcpyDefault, s));
    ELL_CK(cudaStreamSynchronize(s));
    if (seconds) {
        for (int r = 0; r < repeat; r++) {
            float ms = 0.f;
            ELL_CK(cudaEventElapsedTime(&ms, A->events[(size_t)r], A->events[(size_t)r + 1]));
            seconds[r] = (double)ms * 1e-3;
        }
    }
    return 0;
}

int csrspmv_cuda_download(const csrspmv_cuda_matrix *A, int64_t *rowptr, void *colidx, double *a)
{
    if (!A) ELL_FAIL(EINVAL, "matrix is NULL");
    if (!A->shards.empty()) ELL_FAIL(ENOTSUP, "download needs a single-GPU handle");
    if (!rowptr) ELL_FAIL(EINVAL, "rowptr is NULL");
    DeviceGuard g(A->device);
    ELL_CK(cudaMemcpy(rowptr, A->rowptr, (size_t)(A->num_rows + 1) * 8, cudaMemcpyDefault));
    if (A->csrsize > 0) {
        if (!colidx || !a) ELL_FAIL(EINVAL, "colidx or a is NULL");
        ELL_CK(cudaMemcpy(colidx, A->cols, (size_t)A->csrsize * (A->idx_bits / 8), cudaMemcpyDefault));
        ELL_CK(cudaMemcpy(a, A->vals, (size_t)A->csrsize * 8, cudaMemcpyDefault));
    }
    return 0;
}

int64_t csrspmv_cuda_device_bytes(const csrspmv_cuda_matrix *A) { return A ? A->device_bytes : 0; }

int csrspmv_cuda_get_info(const csrspmv_cuda_matrix *A, csrspmv_cuda_info *info)
{
    if (!A || !info) ELL_FAIL(EINVAL, "NULL argument");
    memset(info, 0, sizeof(*info));
    info->num_rows = A->num_rows;
    info->num_columns = A->num_columns;
    info->csrsize = A->csrsize;
    info->device_bytes = A->device_bytes;
    info->fma = A->fma ? 1 : 0;
    info->num_gpus = A->shards.empty() ? 1 : (int)A->shards.size();
    const csrspmv_cuda_matrix *S = A->shards.empty() ? A : A->shards[0];   // a group reports its first shard's kernels
    info->min_row_len = S->min_row_len;
    info->max_row_len = S->max_row_len;
    info->min_col = S->min_col;
    info->max_col = S->max_col;
    info->kernel = S->kernel;
    info->launches_per_spmv = 1;
    if (S->sell) {
        info->sell_slots = sell_entries(S->sell);
        info->sell_real = sell_real_entries(S->sell);
        info->sell_long_rows = sell_long_rows(S->sell);
        info->sell_long_len = sell_long_row(S->sell);
        info->launches_per_spmv = sell_launches(S->sell);
    }
    if (S->ell) {
        info->ell_view = S->ell->d_rowlen ? 1 : 2;
        info->ell_staged = S->ell->sg ? S->ell->staged_mode : 0;
        info->launches_per_spmv = S->ell->sg ? sg_launches(S->ell->sg) : 1;
        info->ell_pattern_rows = S->ell->pat.covered * S->ell->pat.group_rows;
        info->ell_pattern_id_bytes = S->ell->pat.patlane ? S->ell->pat.groups * 32
                                   : (S->ell->pat.max_explicit ? S->ell->pat.groups * 8 : (S->ell->pat.patid ? S->ell->pat.groups : 0));
        info->ell_dev_idx_bits = S->ell->dev_idx_bits;
        info->ell_rows_per_thread = S->ell->cfg.rows_per_thread;
    }
    if (!A->shards.empty()) {
        info->max_row_len = 0;
        info->min_row_len = 0x7fffffffffffffffLL;
        for (const csrspmv_cuda_matrix *T : A->shards) {
            if (T->max_row_len > info->max_row_len) info->max_row_len = T->max_row_len;
            if (T->num_rows > 0 && T->min_row_len < info->min_row_len) info->min_row_len = T->min_row_len;
        }
        if (info->min_row_len == 0x7fffffffffffffffLL) info->min_row_len = 0;
    }
    return 0;
}

/* ---- utilities ------------------------------------------------------------ */

int ellspmv_cuda_malloc_host(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) ELL_FAIL(EINVAL, "bad argument");
    *ptr = nullptr;
    int dev = -1;
    int err = check_device(&dev);
    if (err) return err;
    ELL_CK(cudaMallocHost(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return 0;
}

void ellspmv_cuda_free_host(void *ptr) { if (ptr) cudaFreeHost(ptr); }

int ellspmv_cuda_malloc_device(void **ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) ELL_FAIL(EINVAL, "bad argument");
    *ptr = nullptr;
    int dev = -1;
    int err = check_device(&dev);
    if (err) return err;
    ELL_CK(cudaMalloc(ptr, (size_t)(bytes > 0 ? bytes : 1)));
    return 0;
}

void ellspmv_cuda_free_device(void *ptr) { if (ptr) cudaFree(ptr); }

int ellspmv_cuda_peer_barrier(int rank, int nranks, int64_t epoch, int64_t *local_flags,
                              int64_t *const *peer_flags, void *stream)
{
    if (!local_flags || !peer_flags) ELL_FAIL(EINVAL, "NULL flag array");
    if (nranks < 1 || nranks > kMaxRanks || rank < 0 || rank >= nranks)
        ELL_FAIL(EINVAL, "rank %d / nranks %d out of range (max %d ranks)", rank, nranks, kMaxRanks);
    long long *peers[kMaxRanks];
    for (int p = 0; p < nranks; p++) {
        if (!peer_flags[p]) ELL_FAIL(EINVAL, "peer_flags[%d] is NULL", p);
        peers[p] = reinterpret_cast<long long *>(peer_flags[p]);
    }
    ELL_CK(launch_peer_barrier(rank, nranks, (long long)epoch, reinterpret_cast<long long *>(local_flags), peers,
                               reinterpret_cast<int *>(local_flags + kMaxRanks), (cudaStream_t)stream));
    return 0;
}

int ellspmv_cuda_ipc_export(const void *dev_ptr, unsigned char handle[64])
{
    if (!dev_ptr || !handle) ELL_FAIL(EINVAL, "NULL argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    ELL_CK(cudaIpcGetMemHandle(&h, const_cast<void *>(dev_ptr)));
    memcpy(handle, &h, 64);
    return 0;
}

int ellspmv_cuda_ipc_open(const unsigned char handle[64], void **dev_ptr)
{
    if (!dev_ptr || !handle) ELL_FAIL(EINVAL, "NULL argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    ELL_CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int ellspmv_cuda_ipc_close(void *dev_ptr)
{
    if (!dev_ptr) return 0;
    ELL_CK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

}  // extern "C"

namespace ellspmv {

// rowptr[0] may be non-zero (a row block of a larger matrix): it is rebased on the device
static __global__ void rebase_rowptr_kernel(int64_t *rowptr, int64_t n, int64_t base)
{
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        rowptr[i] -= base;
}

int csr_upload_on(csrspmv_cuda_matrix **out, int idx_width_bits, int64_t num_rows, int64_t num_columns,
                  const int64_t *rowptr, const void *colidx, const double *a, int device, unsigned flags,
                  int64_t row_begin)
{
    if (!rowptr) ELL_FAIL(EINVAL, "rowptr is NULL");
    int err = check_device(&device);
    if (err) return err;
    // rowptr may live on the host or the device: fetch both ends portably
    int64_t first = 0, last = 0;
    ELL_CK(cudaMemcpy(&first, rowptr, 8, cudaMemcpyDefault));
    ELL_CK(cudaMemcpy(&last, rowptr + num_rows, 8, cudaMemcpyDefault));
    const int64_t csrsize = last - first;
    if (csrsize < 0) ELL_FAIL(EINVAL, "rowptr is not increasing");
    if (csrsize > 0 && (!colidx || !a)) ELL_FAIL(EINVAL, "colidx or a is NULL");
    err = csr_new(out, idx_width_bits, num_rows, num_columns, csrsize, device, flags);
    if (err) return err;
    csrspmv_cuda_matrix *A = *out;
    A->row_begin = row_begin;
    DeviceGuard g(A->device);
    const size_t ib = (size_t)idx_width_bits / 8;
    cudaError_t ce = cudaMemcpyAsync(A->rowptr, rowptr, (size_t)(num_rows + 1) * 8, cudaMemcpyDefault, A->stream);
    if (ce == cudaSuccess && first != 0) {
        rebase_rowptr_kernel<<<64, 256, 0, A->stream>>>(A->rowptr, num_rows + 1, first);
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess && csrsize > 0)
        ce = cudaMemcpyAsync(A->cols, (const char *)colidx + (size_t)first * ib, (size_t)csrsize * ib, cudaMemcpyDefault, A->stream);
    if (ce == cudaSuccess && csrsize > 0)
        ce = cudaMemcpyAsync(A->vals, a + first, (size_t)csrsize * 8, cudaMemcpyDefault, A->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(A->stream);
    if (ce != cudaSuccess) {
        set_last_error("csr upload: %s", cudaGetErrorString(ce));
        csrspmv_cuda_free(A); *out = nullptr;
        return cuda_to_errno(ce);
    }
    if ((err = csr_pick_kernel(A))) { csrspmv_cuda_free(A); *out = nullptr; return err; }
    return 0;
}

}  // namespace ellspmv
