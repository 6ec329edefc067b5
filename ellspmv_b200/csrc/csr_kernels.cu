// csr_kernels.cu -- fp64 CSR y <- beta*y + A*x for sm_100a (comparison path).
//
// Replaces the reference's `csrgemv` loop (csrspmv.c:1588-1593):
//     for i: yi = 0; for k in [rowptr[i], rowptr[i+1]): yi += a[k]*x[colidx[k]]; y[i] += yi
//
// csr_stream_kernel (bit-exact): a CTA owns a block of consecutive rows, i.e.
// one contiguous run of entries.  The run is streamed through shared memory
// in tiles: all threads load a[k], colidx[k] with coalesced loads, gather
// x[colidx[k]] and park the rounded product a*x in shared memory; then each
// thread adds up the products of ITS row in entry order.  Loads are coalesced
// no matter how ragged the rows are, and the per-row rounding sequence
// (mul, then left-to-right adds) is the reference's.
//
// csr_vector_kernel (tolerance mode): sub-warp per row with a shuffle tree.
#include "common.cuh"

namespace ellspmv {

constexpr int kCsrRowsPerCta = 128;   // = kBlockThreads: one thread per row in the sum phase
constexpr int kCsrTile = 2048;        // entries staged per pass (16 KB + skew)

// one spare double per 32 breaks the power-of-two stride between rows of
// equal length (K = 32 would otherwise hit one bank from every lane)
__device__ __forceinline__ int skew(int i) { return i + (i >> 5); }

template <typename IdxT, bool FMA>
__global__ void __launch_bounds__(kBlockThreads)
csr_stream_kernel(const CsrSpmvArgs a)
{
    __shared__ double prod[kCsrTile + kCsrTile / 32 + 1];
    const IdxT *__restrict__ cols = reinterpret_cast<const IdxT *>(a.cols);
    const double *__restrict__ vals = a.vals;
    const double *__restrict__ x = a.x;

    const int64_t r0 = (int64_t)blockIdx.x * kCsrRowsPerCta;
    const int64_t r1 = (r0 + kCsrRowsPerCta < a.num_rows) ? r0 + kCsrRowsPerCta : a.num_rows;
    const int64_t row = r0 + threadIdx.x;
    const bool have_row = row < r1;
    const int64_t kb = a.rowptr[r0], ke = a.rowptr[r1];
    int64_t my_b = 0, my_e = 0;
    if (have_row) { my_b = a.rowptr[row]; my_e = a.rowptr[row + 1]; }

    double acc = 0.0;
    for (int64_t t0 = kb; t0 < ke; t0 += kCsrTile) {
        const int64_t t1 = (t0 + kCsrTile < ke) ? t0 + kCsrTile : ke;
        const int n = (int)(t1 - t0);
        // every thread stages kCsrTile/kBlockThreads = 16 entries: all index
        // loads, then all value loads, then all gathers are issued before the
        // first product is parked, so 16 gathers per thread are in flight
        {
            constexpr int PER = kCsrTile / kBlockThreads;
            int64_t c[PER]; double v[PER], xv[PER];
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int i = threadIdx.x + u * kBlockThreads;
                c[u] = i < n ? (int64_t)__ldcs(cols + t0 + i) : 0;
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int i = threadIdx.x + u * kBlockThreads;
                v[u] = i < n ? __ldcs(vals + t0 + i) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int i = threadIdx.x + u * kBlockThreads;
                xv[u] = i < n ? __ldg(x + c[u]) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < PER; u++) {
                const int i = threadIdx.x + u * kBlockThreads;
                if (i < n) prod[skew(i)] = FMA ? v[u] * xv[u] : __dmul_rn(v[u], xv[u]);
            }
        }
        __syncthreads();
        if (have_row) {
            const int64_t b = my_b > t0 ? my_b : t0;
            const int64_t e = my_e < t1 ? my_e : t1;
            for (int64_t k = b; k < e; k++) acc = __dadd_rn(acc, prod[skew((int)(k - t0))]);
        }
        __syncthreads();
    }
    if (have_row) {
        // csrgemvsd (csrspmv.c:1622-1627): y += ad*x + yi
        if (a.ad) acc = __dadd_rn(__dmul_rn(a.ad[row], __ldg(x + a.row_begin + row)), acc);
        const double yold = a.beta ? a.y[row] : 0.0;
        a.y[row] = __dadd_rn(yold, acc);
    }
}

// csr_scalar_kernel (bit-exact): one thread per row walking its entries in
// order.  The loads are not coalesced (neighbouring threads are a row apart),
// but every 32-byte sector a thread touches is used again by its next
// iterations and L1/L2 keep it, so for rows of similar length the DRAM
// traffic is the same as the stream kernel's without its two CTA barriers per
// tile.  Used when the rows are balanced; ragged matrices take the stream kernel.
template <typename IdxT, bool FMA>
__global__ void __launch_bounds__(kBlockThreads)
csr_scalar_kernel(const CsrSpmvArgs a)
{
    const IdxT *__restrict__ cols = reinterpret_cast<const IdxT *>(a.cols);
    const double *__restrict__ vals = a.vals;
    const double *__restrict__ x = a.x;
    const int64_t row = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
    if (row >= a.num_rows) return;
    const int64_t b = a.rowptr[row], e = a.rowptr[row + 1];
    const double yold = a.beta ? a.y[row] : 0.0;
    double acc = 0.0;
    int64_t k = b;
    for (; k + 8 <= e; k += 8) {
        double v[8], xv[8]; int64_t c[8];
#pragma unroll
        for (int u = 0; u < 8; u++) { v[u] = __ldg(vals + k + u); c[u] = (int64_t)__ldg(cols + k + u); }
#pragma unroll
        for (int u = 0; u < 8; u++) xv[u] = __ldg(x + c[u]);
#pragma unroll
        for (int u = 0; u < 8; u++) acc = FMA ? __fma_rn(v[u], xv[u], acc) : __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
    }
    for (; k < e; k++) {
        const double v = __ldg(vals + k);
        const double xv = __ldg(x + (int64_t)__ldg(cols + k));
        acc = FMA ? __fma_rn(v, xv, acc) : __dadd_rn(acc, __dmul_rn(v, xv));
    }
    if (a.ad) acc = __dadd_rn(__dmul_rn(a.ad[row], __ldg(x + a.row_begin + row)), acc);
    a.y[row] = __dadd_rn(yold, acc);
}

// T lanes per row, entries strided by T, shuffle-xor reduction
template <typename IdxT, int T, bool FMA>
__global__ void __launch_bounds__(kBlockThreads)
csr_vector_kernel(const CsrSpmvArgs a)
{
    const IdxT *__restrict__ cols = reinterpret_cast<const IdxT *>(a.cols);
    const double *__restrict__ vals = a.vals;
    const double *__restrict__ x = a.x;
    const int64_t gid = (int64_t)blockIdx.x * kBlockThreads + threadIdx.x;
    const int64_t row = gid / T;
    const int j = (int)(gid % T);
    double acc = 0.0;
    if (row < a.num_rows) {
        const int64_t b = a.rowptr[row], e = a.rowptr[row + 1];
#pragma unroll 4
        for (int64_t k = b + j; k < e; k += T) {
            const double v = __ldcs(vals + k);
            const int64_t c = (int64_t)__ldcs(cols + k);
            const double xv = __ldg(x + c);
            acc = FMA ? __fma_rn(v, xv, acc) : __dadd_rn(acc, __dmul_rn(v, xv));
        }
    }
#pragma unroll
    for (int off = 1; off < T; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (j == 0 && row < a.num_rows) {
        if (a.ad) acc += a.ad[row] * __ldg(x + a.row_begin + row);
        a.y[row] = a.beta ? a.y[row] + acc : acc;
    }
}

// ---- upload-time inspection ----------------------------------------------------
// One pass over rowptr and colidx: the longest row (kernel choice), whether rowptr is
// non-decreasing, and the range of the stored column indices -- the same checks every ELL
// upload makes (finish_minmax in api.cu), so that a bad index is EINVAL at upload and never an
// out-of-bounds gather.  out[0] = max row length, out[1] = rows with rowptr[r+1] < rowptr[r],
// out[2] = min column, out[3] = max column (as signed values), out[4] = min row length.
template <typename IdxT>
__global__ void csr_inspect_kernel(const int64_t *__restrict__ rowptr, const IdxT *__restrict__ cols,
                                   int64_t num_rows, int64_t csrsize, unsigned long long *out)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    long long best = 0, least = 0x7fffffffffffffffLL, bad = 0, lo = 0x7fffffffffffffffLL, hi = -0x7fffffffffffffffLL - 1;
    for (int64_t r = t0; r < num_rows; r += stride) {
        const long long n = rowptr[r + 1] - rowptr[r];
        if (n < 0) bad++;
        best = n > best ? n : best;
        least = n < least ? n : least;
    }
    for (int64_t k = t0; k < csrsize; k += stride) {
        const long long c = (long long)cols[k];
        lo = c < lo ? c : lo;
        hi = c > hi ? c : hi;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const long long ob = __shfl_xor_sync(0xffffffffu, best, off); best = ob > best ? ob : best;
        const long long om = __shfl_xor_sync(0xffffffffu, least, off); least = om < least ? om : least;
        bad += __shfl_xor_sync(0xffffffffu, bad, off);
        const long long ol = __shfl_xor_sync(0xffffffffu, lo, off); lo = ol < lo ? ol : lo;
        const long long oh = __shfl_xor_sync(0xffffffffu, hi, off); hi = oh > hi ? oh : hi;
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(reinterpret_cast<long long *>(out), best);
        if (bad) atomicAdd(out + 1, (unsigned long long)bad);
        atomicMin(reinterpret_cast<long long *>(out) + 2, lo);
        atomicMax(reinterpret_cast<long long *>(out) + 3, hi);
        atomicMin(reinterpret_cast<long long *>(out) + 4, least);
    }
}

// scratch: 5 x 8 bytes of device memory owned by the handle
cudaError_t csr_inspect(int idx_bits, const int64_t *rowptr, const void *cols, int64_t num_rows, int64_t csrsize,
                        unsigned long long *scratch, CsrInspection *res, cudaStream_t stream)
{
    res->max_row_len = 0; res->min_row_len = 0; res->bad_rows = 0; res->min_col = 0; res->max_col = -1;
    if (num_rows <= 0) return cudaSuccess;
    const long long init[5] = {0, 0, 0x7fffffffffffffffLL, -0x7fffffffffffffffLL - 1, 0x7fffffffffffffffLL};
    cudaError_t e = cudaMemcpyAsync(scratch, init, sizeof(init), cudaMemcpyHostToDevice, stream);
    const int64_t work = num_rows > csrsize ? num_rows : csrsize;
    int64_t g = (work + 255) / 256;
    if (g > 148 * 16) g = 148 * 16;
    if (e == cudaSuccess) {
        if (idx_bits == 64)
            csr_inspect_kernel<int64_t><<<(unsigned)g, 256, 0, stream>>>(rowptr, (const int64_t *)cols, num_rows, csrsize, scratch);
        else
            csr_inspect_kernel<int32_t><<<(unsigned)g, 256, 0, stream>>>(rowptr, (const int32_t *)cols, num_rows, csrsize, scratch);
        e = cudaGetLastError();
    }
    long long h[5] = {0, 0, 0, -1, 0};
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, scratch, sizeof(h), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e != cudaSuccess) return e;
    res->max_row_len = h[0];
    res->min_row_len = h[4];
    res->bad_rows = h[1];
    if (csrsize > 0) { res->min_col = h[2]; res->max_col = h[3]; }
    return cudaSuccess;
}

template <typename IdxT, bool FMA>
static cudaError_t launch_csr_typed(int kernel, const CsrSpmvArgs &args, cudaStream_t stream)
{
    if (args.num_rows <= 0) return cudaSuccess;
    if (kernel == ELLSPMV_CUDA_KERNEL_WARP) {
        constexpr int T = 8;
        const int64_t threads = args.num_rows * T;
        const int64_t g = (threads + kBlockThreads - 1) / kBlockThreads;
        if (g > 0x7fffffffLL) return cudaErrorInvalidValue;
        csr_vector_kernel<IdxT, T, FMA><<<(unsigned)g, kBlockThreads, 0, stream>>>(args);
    } else if (kernel == 3) {
        const int64_t g = (args.num_rows + kBlockThreads - 1) / kBlockThreads;
        if (g > 0x7fffffffLL) return cudaErrorInvalidValue;
        csr_scalar_kernel<IdxT, FMA><<<(unsigned)g, kBlockThreads, 0, stream>>>(args);
    } else {
        const int64_t g = (args.num_rows + kCsrRowsPerCta - 1) / kCsrRowsPerCta;
        if (g > 0x7fffffffLL) return cudaErrorInvalidValue;
        csr_stream_kernel<IdxT, FMA><<<(unsigned)g, kBlockThreads, 0, stream>>>(args);
    }
    return cudaGetLastError();
}

cudaError_t launch_csr_spmv(int idx_bits, bool fma, int kernel, const CsrSpmvArgs &args,
                            cudaStream_t stream)
{
    if (idx_bits == 64)
        return fma ? launch_csr_typed<int64_t, true>(kernel, args, stream)
                   : launch_csr_typed<int64_t, false>(kernel, args, stream);
    return fma ? launch_csr_typed<int32_t, true>(kernel, args, stream)
               : launch_csr_typed<int32_t, false>(kernel, args, stream);
}

}  // namespace ellspmv
