// sell.cu -- SELL-128-sigma: sliced ELL with a width PER SLICE, rows sorted by length inside
// windows of sigma rows, every row summed by one thread in slot order.  Bit-exact.
//
// Why: the reference pads every row to the longest one (ellspmv.c:944-955, 1111-1117) and its CSR
// loop gives a thread whole rows whatever their length (csrspmv.c:1588-1593; csrgemvnz,
// csrspmv.c:1698-1740, balances entries between THREADS with atomics).  On the GPU a padded slot
// is HBM traffic, and a warp runs as long as its longest row.  Here
//   * rows are sorted by length (descending) inside windows of 4096 rows, so the 128 rows of a
//     slice -- and the 32 of a warp -- have nearly equal lengths;
//   * a slice stores only width[s] = its longest row's slots, slot-major (element (row p, slot l)
//     at slice_ptr[s] + l*128 + p%128): the loads are the coalesced streams of the ELL kernel,
//     the padding left is the spread of lengths inside one slice;
//   * each thread adds ITS row's products, slots 0..len-1, mul then add: the reference's
//     rounding sequence; slots past a row's end are never touched arithmetically, so there is no
//     0*inf from padding and the result equals csrgemv's bits also for non-finite x;
//   * y is read and written through the row permutation (scattered inside one 32 KB window);
//   * rows longer than 256 entries (a power-law tail) would still set the width of their slice
//     (128 rows x the longest one) and serialise a warp: they are left out of the slices and run
//     one CTA per row -- all threads streaming + gathering, rounded products parked in shared
//     memory, one thread adding them in order (the scheme of ell_longrow.cu), reading the CSR
//     arrays in place -- on a second stream NEXT TO the slice kernel: a row of 200 000 entries is
//     a 1.4 ms chain of dependent additions whoever runs it, and it now hides behind the slices.
//
// Used by the CSR path (KERNEL_AUTO with unbalanced rows, api.cu: csr_build_sell) and, opt-in, by
// the ELL path (ELLSPMV_CUDA_SKIP_PADDING: a row's trailing reference padding -- column
// min(i, ncols-1), value 0.0 -- is cut off; exact for finite x, differs where x is non-finite on a
// padding column, which is why it is opt-in there).
#include <stdlib.h>

#include <cub/cub.cuh>

#include "common.cuh"

namespace ellspmv {

constexpr int kSellSlice = 128;              // rows per slice = threads per CTA
constexpr int kSellWindow = 4096;            // sigma: rows sorted by length inside windows of this size
constexpr int kSellLongRowDefault = 256;     // CSR rows longer than this run one CTA per row (SELL_LONG_ROW overrides)

struct SellMatrix {
    int idx_bits = 32;
    int64_t num_rows = 0, padded_rows = 0, num_slices = 0;
    int64_t entries = 0;                     // stored slots (sum of width * 128)
    int64_t real_entries = 0;                // entries that count (sum of the row lengths in the slices)
    long long *slice_ptr = nullptr;          // num_slices + 1
    int *rowlen = nullptr;                   // per packed row: slots that count (0: empty or padding row; -1: a long row)
    int *perm = nullptr;                     // packed row -> row of the matrix (-1: padding)
    void *cols = nullptr;
    double *vals = nullptr;
    int64_t num_long = 0;
    long long *long_rows = nullptr;          // rows run by the CTA-per-row kernel
    double *long_sum = nullptr;              // their row sums, parked between the kernels of one SpMV
    int long_row = 0;                        // the length from which a row counts as long (0: none do)
    cudaStream_t side = nullptr;             // the long rows run here, next to the slices on the caller's stream
    cudaEvent_t fork = nullptr, join = nullptr;
    int64_t bytes = 0;
};

void sell_free(SellMatrix *m)
{
    if (!m) return;
    cudaFree(m->slice_ptr); cudaFree(m->rowlen); cudaFree(m->perm); cudaFree(m->cols); cudaFree(m->vals);
    cudaFree(m->long_rows); cudaFree(m->long_sum);
    if (m->side) cudaStreamDestroy(m->side);
    if (m->fork) cudaEventDestroy(m->fork);
    if (m->join) cudaEventDestroy(m->join);
    delete m;
}
int64_t sell_bytes(const SellMatrix *m) { return m ? m->bytes : 0; }
int64_t sell_entries(const SellMatrix *m) { return m ? m->entries : 0; }
int64_t sell_real_entries(const SellMatrix *m) { return m ? m->real_entries : 0; }
int64_t sell_long_rows(const SellMatrix *m) { return m ? m->num_long : 0; }
int sell_launches(const SellMatrix *m) { return m ? 1 + (m->num_long > 0 ? 2 : 0) : 0; }
int sell_long_row(const SellMatrix *m) { return m ? m->long_row : 0; }

// ---- where a row's entries live ----------------------------------------------------------------
struct CsrSource {                             // CSR arrays: row i at [rowptr[i], rowptr[i+1])
    const int64_t *rowptr;
    __device__ __forceinline__ int64_t length(int64_t i) const { return rowptr[i + 1] - rowptr[i]; }
    __device__ __forceinline__ int64_t at(int64_t i, int l) const { return rowptr[i] + l; }
};
struct EllSource {                             // sliced ELL arrays of a handle, lengths from ell_trim_kernel
    EllLayout lay;
    const int *len;
    __device__ __forceinline__ int64_t length(int64_t i) const { return len[i]; }
    __device__ __forceinline__ int64_t at(int64_t i, int l) const { return lay.offset(i, l); }
};

// a row's length without its trailing reference padding (col == min(i, ncols-1) && val == 0.0)
template <typename IdxT>
__global__ void ell_trim_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, EllLayout lay,
                                int64_t row_begin, int64_t num_columns, int *__restrict__ len)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= lay.num_rows) return;
    const int64_t g = row_begin + i;
    const int64_t pad = g < num_columns ? g : num_columns - 1;
    int n = lay.rowsize;
    while (n > 0) {
        const int64_t o = lay.offset(i, n - 1);
        if ((int64_t)cols[o] == pad && vals[o] == 0.0) n--; else break;
    }
    len[i] = n;
}

// ---- build ----------------------------------------------------------------------------------------
// one CTA per window: sort (length, row) descending by length -- bitonic in shared memory -- and
// write the permutation, the packed lengths and the widths of the window's slices
template <typename Src>
__global__ void __launch_bounds__(1024)
sell_sort_kernel(Src src, int64_t num_rows, int long_row, int *__restrict__ perm, int *__restrict__ rowlen,
                 long long *__restrict__ slots /* per slice: width * 128 */, unsigned long long *__restrict__ stats)
{
    __shared__ unsigned long long key[kSellWindow];
    const int64_t w0 = (int64_t)blockIdx.x * kSellWindow;
    unsigned long long nlong = 0, real = 0;
    for (int i = threadIdx.x; i < kSellWindow; i += blockDim.x) {
        const int64_t row = w0 + i;
        unsigned long long k = 0;                                  // padding rows: length 0, sorted last
        if (row < num_rows) {
            int64_t len = src.length(row);
            unsigned long long is_long = 0;
            if (long_row > 0 && len > long_row) { len = 0; is_long = 1; nlong++; }   // runs in the CTA-per-row kernel
            real += (unsigned long long)len;
            // descending by length, ascending by row among equals:
            // key = len << 32 | long-row mark << 16 | (4095 - position in the window)
            k = ((unsigned long long)len << 32) | (is_long << 16) | (unsigned long long)(kSellWindow - 1 - i);
        }
        key[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= kSellWindow; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = threadIdx.x; i < kSellWindow / 2; i += blockDim.x) {
                const int lo = 2 * i - (i & (stride - 1));        // i-th pair of this stage
                const int hi = lo + stride;
                const bool desc = (lo & size) == 0;
                const unsigned long long a = key[lo], b = key[hi];
                if ((a < b) == desc) { key[lo] = b; key[hi] = a; }
            }
            __syncthreads();
        }
    for (int i = threadIdx.x; i < kSellWindow; i += blockDim.x) {
        const unsigned long long k = key[i];
        const int wi = kSellWindow - 1 - (int)(k & 0xfffu);       // the row's position in the window
        const int64_t row = w0 + wi;
        const int len = (int)(k >> 32);
        const int64_t p = w0 + i;
        perm[p] = row < num_rows ? wi : -1;                       // window-relative (fits an int at any size)
        rowlen[p] = row < num_rows ? (((k >> 16) & 1ull) ? -1 : len) : 0;    // -1: a long row, not ours
        if ((i & (kSellSlice - 1)) == 0) slots[p / kSellSlice] = (long long)len * kSellSlice;   // the slice's longest row
    }
    for (int off = 16; off > 0; off >>= 1) {
        nlong += __shfl_xor_sync(0xffffffffu, nlong, off);
        real += __shfl_xor_sync(0xffffffffu, real, off);
    }
    if ((threadIdx.x & 31) == 0) { if (nlong) atomicAdd(stats, nlong); if (real) atomicAdd(stats + 1, real); }
}

// one CTA per slice: copy the rows' entries into the slot-major slice, fill what is left of the
// slice's width with (column 0, 0.0) -- never used arithmetically
template <typename Src, typename SrcI, typename DstI>
__global__ void __launch_bounds__(kSellSlice)
sell_fill_kernel(Src src, const SrcI *__restrict__ src_cols, const double *__restrict__ src_vals,
                 const int *__restrict__ perm, const int *__restrict__ rowlen, const long long *__restrict__ slice_ptr,
                 DstI *__restrict__ cols, double *__restrict__ vals)
{
    const int64_t s = blockIdx.x;
    const int64_t base = slice_ptr[s];
    const int width = (int)((slice_ptr[s + 1] - base) / kSellSlice);
    const int64_t p = s * kSellSlice + threadIdx.x;
    const int len = rowlen[p];                 // <= 0: nothing of this row is stored here
    const int64_t row = (p / kSellWindow) * kSellWindow + perm[p];
    for (int l = 0; l < width; l++) {
        const int64_t d = base + (int64_t)l * kSellSlice + threadIdx.x;
        if (l < len) {
            const int64_t o = src.at(row, l);
            cols[d] = (DstI)src_cols[o];
            vals[d] = src_vals[o];
        } else {
            cols[d] = (DstI)0;
            vals[d] = 0.0;
        }
    }
}

template <typename Src>
__global__ void sell_long_list_kernel(Src src, int64_t num_rows, int long_row, long long *__restrict__ list,
                                      long long *__restrict__ lens, unsigned long long *__restrict__ cursor)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= num_rows) return;
    const int64_t n = src.length(i);
    if (n > long_row) {
        const unsigned long long at = atomicAdd(cursor, 1ull);
        list[at] = i;
        lens[at] = n;
    }
}

template <typename Src, typename SrcI>
static cudaError_t sell_build_typed(SellMatrix *m, Src src, const SrcI *src_cols, const double *src_vals,
                                    int dst_idx_bits, int long_row, cudaStream_t stream)
{
    const int64_t n = m->num_rows;
    const int64_t windows = (n + kSellWindow - 1) / kSellWindow;
    m->padded_rows = windows * kSellWindow;
    m->num_slices = m->padded_rows / kSellSlice;
    cudaError_t e;
    long long *slots = nullptr;
    unsigned long long *stats = nullptr;
    void *temp = nullptr;
    auto cleanup = [&]() { cudaFree(slots); cudaFree(stats); cudaFree(temp); };
    if ((e = cudaMalloc(&m->perm, (size_t)m->padded_rows * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&m->rowlen, (size_t)m->padded_rows * sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&m->slice_ptr, (size_t)(m->num_slices + 1) * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&slots, (size_t)(m->num_slices + 1) * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&stats, 24)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMemsetAsync(stats, 0, 24, stream)) != cudaSuccess ||
        (e = cudaMemsetAsync(slots, 0, (size_t)(m->num_slices + 1) * 8, stream)) != cudaSuccess) { cleanup(); return e; }
    sell_sort_kernel<Src><<<(unsigned)windows, 1024, 0, stream>>>(src, n, long_row, m->perm, m->rowlen, slots, stats);
    if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
    size_t tb = 0;
    e = cub::DeviceScan::ExclusiveSum(nullptr, tb, slots, m->slice_ptr, m->num_slices + 1, stream);
    if (e == cudaSuccess) e = cudaMalloc(&temp, tb + 16);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(temp, tb, slots, m->slice_ptr, m->num_slices + 1, stream);
    unsigned long long hs[2] = {0, 0};
    long long total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(hs, stats, 16, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, m->slice_ptr + m->num_slices, 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) { cleanup(); return e; }
    m->num_long = (int64_t)hs[0];
    m->real_entries = (int64_t)hs[1];
    m->entries = total;
    const size_t ne = (size_t)(total > 0 ? total : 1);
    const size_t ib = (size_t)dst_idx_bits / 8;
    if ((e = cudaMalloc(&m->cols, ne * ib)) != cudaSuccess) { cleanup(); return e; }
    if ((e = cudaMalloc(&m->vals, ne * 8)) != cudaSuccess) { cleanup(); return e; }
    if (total > 0) {
        if (dst_idx_bits == 64)
            sell_fill_kernel<Src, SrcI, int64_t><<<(unsigned)m->num_slices, kSellSlice, 0, stream>>>(
                src, src_cols, src_vals, m->perm, m->rowlen, m->slice_ptr, (int64_t *)m->cols, m->vals);
        else
            sell_fill_kernel<Src, SrcI, int32_t><<<(unsigned)m->num_slices, kSellSlice, 0, stream>>>(
                src, src_cols, src_vals, m->perm, m->rowlen, m->slice_ptr, (int32_t *)m->cols, m->vals);
        if ((e = cudaGetLastError()) != cudaSuccess) { cleanup(); return e; }
    }
    if (m->num_long > 0) {
        if ((e = cudaMalloc(&m->long_rows, (size_t)m->num_long * 8)) != cudaSuccess) { cleanup(); return e; }
        if ((e = cudaMalloc(&m->long_sum, (size_t)m->num_long * 8)) != cudaSuccess) { cleanup(); return e; }
        if ((e = cudaMemsetAsync(stats + 2, 0, 8, stream)) != cudaSuccess) { cleanup(); return e; }
        // the list, longest row first: a row is one chain of dependent additions, so the longest one
        // sets the finish time and must start first (stable sort of (length, row) by length, descending)
        long long *rows_unsorted = nullptr, *lens = nullptr, *lens_sorted = nullptr;
        void *stemp = nullptr;
        size_t sb = 0;
        auto cleanup2 = [&]() { cudaFree(rows_unsorted); cudaFree(lens); cudaFree(lens_sorted); cudaFree(stemp); };
        if ((e = cudaMalloc(&rows_unsorted, (size_t)m->num_long * 8)) != cudaSuccess ||
            (e = cudaMalloc(&lens, (size_t)m->num_long * 8)) != cudaSuccess ||
            (e = cudaMalloc(&lens_sorted, (size_t)m->num_long * 8)) != cudaSuccess) { cleanup2(); cleanup(); return e; }
        sell_long_list_kernel<Src><<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(src, n, long_row, rows_unsorted, lens, stats + 2);
        e = cudaGetLastError();
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairsDescending(nullptr, sb, lens, lens_sorted, rows_unsorted, m->long_rows,
                                                          m->num_long, 0, 64, stream);
        if (e == cudaSuccess) e = cudaMalloc(&stemp, sb + 16);
        if (e == cudaSuccess)
            e = cub::DeviceRadixSort::SortPairsDescending(stemp, sb, lens, lens_sorted, rows_unsorted, m->long_rows,
                                                          m->num_long, 0, 64, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        cleanup2();
        if (e != cudaSuccess) { cleanup(); return e; }
        if ((e = cudaStreamCreateWithFlags(&m->side, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&m->fork, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&m->join, cudaEventDisableTiming)) != cudaSuccess) { cleanup(); return e; }
    }
    e = cudaStreamSynchronize(stream);
    cleanup();
    m->idx_bits = dst_idx_bits;
    m->long_row = long_row;
    m->bytes = (int64_t)(ne * (ib + 8)) + m->padded_rows * 8 + (m->num_slices + 1) * 8 + m->num_long * 16;
    return e;
}

// from CSR arrays on the device (rows longer than kSellLongRow go to the CTA-per-row kernel)
cudaError_t sell_build_csr(SellMatrix **out, int src_idx_bits, int dst_idx_bits, int64_t num_rows, const int64_t *rowptr,
                           const void *cols, const double *vals, cudaStream_t stream)
{
    *out = nullptr;
    if (num_rows <= 0) return cudaSuccess;
    if ((num_rows + kSellWindow - 1) / kSellWindow > 0x7fffffffLL) return cudaErrorInvalidValue;
    SellMatrix *m = new (std::nothrow) SellMatrix();
    if (!m) return cudaErrorMemoryAllocation;
    m->num_rows = num_rows;
    CsrSource src{rowptr};
    int long_row = kSellLongRowDefault;
    if (const char *env = getenv("SELL_LONG_ROW")) { const int v = atoi(env); if (v >= 32) long_row = v; }
    cudaError_t e = src_idx_bits == 64
        ? sell_build_typed<CsrSource, int64_t>(m, src, (const int64_t *)cols, vals, dst_idx_bits, long_row, stream)
        : sell_build_typed<CsrSource, int32_t>(m, src, (const int32_t *)cols, vals, dst_idx_bits, long_row, stream);
    if (e != cudaSuccess) { sell_free(m); return e; }
    *out = m;
    return cudaSuccess;
}

// from a handle's sliced ELL arrays, every row cut after its last slot that is not reference padding
cudaError_t sell_build_ell(SellMatrix **out, int idx_bits, const void *cols, const double *vals, const EllLayout &lay,
                           int64_t row_begin, int64_t num_columns, cudaStream_t stream)
{
    *out = nullptr;
    if (lay.num_rows <= 0 || lay.rowsize <= 0) return cudaSuccess;
    SellMatrix *m = new (std::nothrow) SellMatrix();
    if (!m) return cudaErrorMemoryAllocation;
    m->num_rows = lay.num_rows;
    int *len = nullptr;
    cudaError_t e = cudaMalloc(&len, (size_t)lay.num_rows * sizeof(int));
    if (e != cudaSuccess) { sell_free(m); return e; }
    const unsigned g = (unsigned)((lay.num_rows + 255) / 256);
    if (idx_bits == 64) ell_trim_kernel<int64_t><<<g, 256, 0, stream>>>((const int64_t *)cols, vals, lay, row_begin, num_columns, len);
    else ell_trim_kernel<int32_t><<<g, 256, 0, stream>>>((const int32_t *)cols, vals, lay, row_begin, num_columns, len);
    e = cudaGetLastError();
    EllSource src{lay, len};
    if (e == cudaSuccess)
        e = idx_bits == 64 ? sell_build_typed<EllSource, int64_t>(m, src, (const int64_t *)cols, vals, idx_bits, 0, stream)
                           : sell_build_typed<EllSource, int32_t>(m, src, (const int32_t *)cols, vals, idx_bits, 0, stream);
    cudaFree(len);
    if (e != cudaSuccess) { sell_free(m); return e; }
    *out = m;
    return cudaSuccess;
}

// ---- the kernel: one thread per (packed) row, slots 0..len-1 in order --------------------------
template <typename IdxT, bool FMA>
__global__ void __launch_bounds__(kSellSlice)
sell_spmv_kernel(const double *__restrict__ vals, const IdxT *__restrict__ cols, const long long *__restrict__ slice_ptr,
                 const int *__restrict__ rowlen, const int *__restrict__ perm, const double *__restrict__ x,
                 double *__restrict__ y, const double *__restrict__ ad, int64_t row_begin, int beta)
{
    constexpr int U = 8;
    const int64_t s = blockIdx.x;
    const int64_t p = s * kSellSlice + threadIdx.x;
    const int pr = perm[p];
    const int len_raw = rowlen[p];
    const int len = len_raw > 0 ? len_raw : 0;
    const int64_t row = (p / kSellWindow) * kSellWindow + pr;
    const int wmax = __reduce_max_sync(0xffffffffu, len);           // the warp stops at its longest row
    const double *vp = vals + slice_ptr[s] + threadIdx.x;
    const IdxT *cp = cols + slice_ptr[s] + threadIdx.x;
    double acc = 0.0;
    int l0 = 0;
#pragma unroll 1
    for (; l0 + U <= wmax; l0 += U) {
        double v[U], xv[U]; int64_t c[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v[u]) : "l"(vp + (int64_t)(l0 + u) * kSellSlice));
            c[u] = (int64_t)__ldcs(cp + (int64_t)(l0 + u) * kSellSlice);
        }
#pragma unroll
        for (int u = 0; u < U; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
        for (int u = 0; u < U; u++)
            if (l0 + u < len) acc = FMA ? __fma_rn(v[u], xv[u], acc) : __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
    }
#pragma unroll 1
    for (; l0 < wmax; l0++) {
        double v;
        asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(vp + (int64_t)l0 * kSellSlice));
        const int64_t c = (int64_t)__ldcs(cp + (int64_t)l0 * kSellSlice);
        const double xv = __ldg(x + c);
        if (l0 < len) acc = FMA ? __fma_rn(v, xv, acc) : __dadd_rn(acc, __dmul_rn(v, xv));
    }
    // padding rows of the last window have no y; a long row (-1) is finished by sell_long_finish_kernel
    if (pr < 0 || len_raw < 0) return;
    // csrgemvsd (csrspmv.c:1622-1627) / ellgemvsd order 0: ad*x + yi
    if (ad) acc = __dadd_rn(__dmul_rn(ad[row], __ldg(x + row_begin + row)), acc);
    const double yold = beta ? y[row] : 0.0;
    y[row] = __dadd_rn(yold, acc);
}

// rows of the long list: one CTA each, everything streamed by all threads, the rounded products
// parked in shared memory, thread 0 adds them in order (the scheme of ell_longrow.cu, reading the
// CSR arrays in place).  The row sums are parked in long_sum; sell_long_finish_kernel applies
// them to y after the slice kernel, which leaves the y entries of these rows alone.
// adds n parked products to acc in order; the next 8 operands are fetched from shared memory
// while the current 8 are being added, so the chain runs at the DADD latency, not DADD + LDS
__device__ __forceinline__ double sum_in_order(double acc, const double *q, int n)
{
    int l = 0;
    if (n >= 8) {
        double w[8];
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = q[j];
        for (l = 8; l + 8 <= n; l += 8) {
            double wn[8];
#pragma unroll
            for (int j = 0; j < 8; j++) wn[j] = q[l + j];
#pragma unroll
            for (int j = 0; j < 8; j++) acc = __dadd_rn(acc, w[j]);
#pragma unroll
            for (int j = 0; j < 8; j++) w[j] = wn[j];
        }
#pragma unroll
        for (int j = 0; j < 8; j++) acc = __dadd_rn(acc, w[j]);
    }
    for (; l < n; l++) acc = __dadd_rn(acc, q[l]);
    return acc;
}

template <typename IdxT>
__global__ void __launch_bounds__(128)
sell_long_kernel(const int64_t *__restrict__ rowptr, const IdxT *__restrict__ cols, const double *__restrict__ vals,
                 const long long *__restrict__ list, const double *__restrict__ x, double *__restrict__ long_sum)
{
    constexpr int NT = 128, T = 1024, PER = T / NT;
    __shared__ double prod[2][T];
    const int64_t row = list[blockIdx.x];
    const int64_t kb = rowptr[row], ke = rowptr[row + 1];
    const int tid = threadIdx.x;
    const int64_t ntiles = (ke - kb + T - 1) / T;
    auto issue_vc = [&](int64_t t, double (&v)[PER], int64_t (&c)[PER]) {
#pragma unroll
        for (int u = 0; u < PER; u++) {
            const int64_t k = kb + t * T + u * NT + tid;
            v[u] = 0.0; c[u] = -1;
            if (t < ntiles && k < ke) { v[u] = __ldcs(vals + k); c[u] = (int64_t)__ldcs(cols + k); }
        }
    };
    double v1[PER], x1[PER], v2[PER]; int64_t c2[PER];
    {
        int64_t c1[PER];
        issue_vc(0, v1, c1);
        issue_vc(1, v2, c2);
#pragma unroll
        for (int u = 0; u < PER; u++) x1[u] = c1[u] >= 0 ? __ldg(x + c1[u]) : 0.0;
    }
    double acc = 0.0;
    for (int64_t t = 0; t < ntiles; t++) {
        double *p = prod[t & 1];
#pragma unroll
        for (int u = 0; u < PER; u++) p[u * NT + tid] = __dmul_rn(v1[u], x1[u]);
        __syncthreads();
#pragma unroll
        for (int u = 0; u < PER; u++) { v1[u] = v2[u]; x1[u] = c2[u] >= 0 ? __ldg(x + c2[u]) : 0.0; }
        issue_vc(t + 2, v2, c2);
        if (tid == 0) {
            const int64_t left = ke - kb - t * T;
            acc = sum_in_order(acc, p, left < T ? (int)left : T);
        }
    }
    if (tid == 0) long_sum[blockIdx.x] = acc;
}

// applies the parked sums of the long rows: y[row] = yold + (ad*x + sum), the same roundings as a
// row of the slice kernel (which left these rows alone: rowlen == -1)
__global__ void sell_long_finish_kernel(const long long *__restrict__ list, const double *__restrict__ long_sum, int64_t n,
                                        const double *__restrict__ x, double *__restrict__ y,
                                        const double *__restrict__ ad, int64_t row_begin, int beta)
{
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t row = list[i];
    double acc = long_sum[i];
    if (ad) acc = __dadd_rn(__dmul_rn(ad[row], __ldg(x + row_begin + row)), acc);
    const double yold = beta ? y[row] : 0.0;
    y[row] = __dadd_rn(yold, acc);
}

// y <- beta*y + A*x: the long rows' sums (they need x only), the slices, then the long rows' y entries
cudaError_t sell_spmv(const SellMatrix *m, bool fma, const int64_t *csr_rowptr, const void *csr_cols,
                      const double *csr_vals, int csr_idx_bits, const double *x, double *y,
                      const double *ad, int64_t row_begin, int beta, cudaStream_t stream)
{
    if (!m || m->num_rows <= 0) return cudaSuccess;
    if (m->num_long > 0) {
        if (!csr_rowptr) return cudaErrorInvalidValue;
        // fork: the long rows need x only; they run on the side stream while the slices run here
        cudaError_t e = cudaEventRecord(m->fork, stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(m->side, m->fork, 0);
        if (e != cudaSuccess) return e;
        if (csr_idx_bits == 64)
            sell_long_kernel<int64_t><<<(unsigned)m->num_long, 128, 0, m->side>>>(csr_rowptr, (const int64_t *)csr_cols, csr_vals, m->long_rows, x, m->long_sum);
        else
            sell_long_kernel<int32_t><<<(unsigned)m->num_long, 128, 0, m->side>>>(csr_rowptr, (const int32_t *)csr_cols, csr_vals, m->long_rows, x, m->long_sum);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaEventRecord(m->join, m->side);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)m->num_slices;
    if (m->idx_bits == 64) {
        auto k = fma ? sell_spmv_kernel<int64_t, true> : sell_spmv_kernel<int64_t, false>;
        k<<<grid, kSellSlice, 0, stream>>>(m->vals, (const int64_t *)m->cols, m->slice_ptr, m->rowlen, m->perm, x, y, ad, row_begin, beta);
    } else {
        auto k = fma ? sell_spmv_kernel<int32_t, true> : sell_spmv_kernel<int32_t, false>;
        k<<<grid, kSellSlice, 0, stream>>>(m->vals, (const int32_t *)m->cols, m->slice_ptr, m->rowlen, m->perm, x, y, ad, row_begin, beta);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if (m->num_long > 0) {
        if ((e = cudaStreamWaitEvent(stream, m->join, 0)) != cudaSuccess) return e;      // join
        sell_long_finish_kernel<<<(unsigned)((m->num_long + 127) / 128), 128, 0, stream>>>(m->long_rows, m->long_sum, m->num_long, x, y, ad, row_begin, beta);
        e = cudaGetLastError();
    }
    return e;
}

}  // namespace ellspmv
