// layout.cu -- device re-layout of the reference's row-major ELL arrays into
// sliced ELL, its inverse, and on-device generators of BASELINE.json's
// synthetic matrices.
//
// The row-major arrays are what the reference's ell_from_coo produces
// (ellspmv.c:1081-1127): colidx[i*K+l], a[i*K+l], 0-based, rows padded with
// (min(i, ncols-1), 0.0).  Re-layout is a pure permutation (plus optional
// 64->32-bit index narrowing), so download(upload(A)) == A bit for bit.
#include "common.cuh"

namespace ellspmv {

__device__ __forceinline__ void block_minmax(long long lo, long long hi, long long *minmax)
{
    // warp reduce, then one atomic pair per warp
    for (int off = 16; off > 0; off >>= 1) {
        long long olo = __shfl_xor_sync(0xffffffffu, lo, off);
        long long ohi = __shfl_xor_sync(0xffffffffu, hi, off);
        lo = olo < lo ? olo : lo;
        hi = ohi > hi ? ohi : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) {
        atomicMin(minmax, lo);
        atomicMax(minmax + 1, hi);
    }
}

__global__ void init_minmax_kernel(long long *minmax)
{
    minmax[0] = 0x7fffffffffffffffLL;
    minmax[1] = -1;
}

cudaError_t init_minmax(long long *minmax, cudaStream_t stream)
{
    init_minmax_kernel<<<1, 1, 0, stream>>>(minmax);
    return cudaGetLastError();
}

// one thread per (row, slot) element of the row-major chunk: coalesced read,
// scattered (but slice-local) write
template <typename SrcI, typename DstI>
__global__ void relayout_kernel(const SrcI *__restrict__ src_cols, const double *__restrict__ src_vals,
                                DstI *__restrict__ dst_cols, double *__restrict__ dst_vals,
                                EllLayout lay, int64_t row0, int64_t rows, long long *minmax)
{
    const int K = lay.rowsize;
    const int64_t n = rows * K;
    long long lo = 0x7fffffffffffffffLL, hi = -1;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / K;
        const int l = (int)(e - r * K);
        const long long c = (long long)src_cols[e];
        const int64_t d = lay.offset(row0 + r, l);
        dst_cols[d] = (DstI)c;
        dst_vals[d] = src_vals[e];
        lo = c < lo ? c : lo;
        hi = c > hi ? c : hi;
    }
    if (minmax) block_minmax(lo, hi, minmax);
}

template <typename SrcI, typename DstI>
__global__ void unlayout_kernel(const SrcI *__restrict__ src_cols, const double *__restrict__ src_vals,
                                DstI *__restrict__ dst_cols, double *__restrict__ dst_vals,
                                EllLayout lay, int64_t row0, int64_t rows)
{
    const int K = lay.rowsize;
    const int64_t n = rows * K;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = e / K;
        const int l = (int)(e - r * K);
        const int64_t s = lay.offset(row0 + r, l);
        dst_cols[e] = (DstI)src_cols[s];
        dst_vals[e] = src_vals[s];
    }
}

static int grid_for(int64_t n, int block)
{
    int64_t g = (n + block - 1) / block;
    if (g > 148 * 64) g = 148 * 64;
    if (g < 1) g = 1;
    return (int)g;
}

cudaError_t relayout_chunk(int src_idx_bits, int dst_idx_bits, const void *src_cols,
                           const double *src_vals, void *dst_cols, double *dst_vals,
                           const EllLayout &lay, int64_t row0, int64_t rows,
                           long long *minmax, cudaStream_t stream)
{
    if (rows <= 0 || lay.rowsize <= 0) return cudaSuccess;
    const int64_t n = rows * lay.rowsize;
    const int g = grid_for(n, 256);
    if (src_idx_bits == 32 && dst_idx_bits == 32)
        relayout_kernel<int32_t, int32_t><<<g, 256, 0, stream>>>((const int32_t *)src_cols, src_vals, (int32_t *)dst_cols, dst_vals, lay, row0, rows, minmax);
    else if (src_idx_bits == 64 && dst_idx_bits == 64)
        relayout_kernel<int64_t, int64_t><<<g, 256, 0, stream>>>((const int64_t *)src_cols, src_vals, (int64_t *)dst_cols, dst_vals, lay, row0, rows, minmax);
    else if (src_idx_bits == 64 && dst_idx_bits == 32)
        relayout_kernel<int64_t, int32_t><<<g, 256, 0, stream>>>((const int64_t *)src_cols, src_vals, (int32_t *)dst_cols, dst_vals, lay, row0, rows, minmax);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

cudaError_t unlayout_chunk(int dev_idx_bits, int host_idx_bits, const void *src_cols,
                           const double *src_vals, void *dst_cols, double *dst_vals,
                           const EllLayout &lay, int64_t row0, int64_t rows,
                           cudaStream_t stream)
{
    if (rows <= 0 || lay.rowsize <= 0) return cudaSuccess;
    const int64_t n = rows * lay.rowsize;
    const int g = grid_for(n, 256);
    if (dev_idx_bits == 32 && host_idx_bits == 32)
        unlayout_kernel<int32_t, int32_t><<<g, 256, 0, stream>>>((const int32_t *)src_cols, src_vals, (int32_t *)dst_cols, dst_vals, lay, row0, rows);
    else if (dev_idx_bits == 64 && host_idx_bits == 64)
        unlayout_kernel<int64_t, int64_t><<<g, 256, 0, stream>>>((const int64_t *)src_cols, src_vals, (int64_t *)dst_cols, dst_vals, lay, row0, rows);
    else if (dev_idx_bits == 32 && host_idx_bits == 64)
        unlayout_kernel<int32_t, int64_t><<<g, 256, 0, stream>>>((const int32_t *)src_cols, src_vals, (int64_t *)dst_cols, dst_vals, lay, row0, rows);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// ---- synthetic generators -----------------------------------
// One thread per shard row; for each slot the threads of a CTA write
// consecutive addresses of the sliced layout.  Entries, order and padding are
// those ell_from_coo (ellspmv.c:1081-1127) would produce from the canonical
// COO stream; tests compare against that route.

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9e3779b97f4a7c15ULL;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}

template <typename DstI>
__global__ void gen_laplace2d_kernel(int64_t nx, int64_t ny, double cval, double oval,
                                     DstI *__restrict__ cols, double *__restrict__ vals,
                                     EllLayout lay, int64_t row_begin, long long *minmax)
{
    const int64_t ncols = nx * ny;
    long long lo = 0x7fffffffffffffffLL, hi = -1;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < lay.num_rows;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_begin + q;
        const int64_t i = r / ny, j = r - i * ny;
        int64_t c[5]; double v[5]; int n = 0;
        if (i > 0)      { c[n] = r - ny; v[n] = oval; n++; }
        if (j > 0)      { c[n] = r - 1;  v[n] = oval; n++; }
        c[n] = r; v[n] = cval; n++;
        if (j + 1 < ny) { c[n] = r + 1;  v[n] = oval; n++; }
        if (i + 1 < nx) { c[n] = r + ny; v[n] = oval; n++; }
        const int64_t pad = r < ncols ? r : ncols - 1;
        for (; n < 5; n++) { c[n] = pad; v[n] = 0.0; }
#pragma unroll
        for (int l = 0; l < 5; l++) {
            const int64_t d = lay.offset(q, l);
            cols[d] = (DstI)c[l]; vals[d] = v[l];
            lo = c[l] < lo ? c[l] : lo; hi = c[l] > hi ? c[l] : hi;
        }
    }
    if (minmax) block_minmax(lo, hi, minmax);
}

template <typename DstI>
__global__ void gen_stencil27_kernel(int64_t nx, int64_t ny, int64_t nz, double cval, double oval,
                                     DstI *__restrict__ cols, double *__restrict__ vals,
                                     EllLayout lay, int64_t row_begin, long long *minmax)
{
    const int64_t ncols = nx * ny * nz;
    long long lo = 0x7fffffffffffffffLL, hi = -1;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < lay.num_rows;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_begin + q;
        const int64_t k = r % nz, j = (r / nz) % ny, i = r / (nz * ny);
        int n = 0;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++) {
                    const int64_t ii = i + di, jj = j + dj, kk = k + dk;
                    if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
                    const int64_t c = (ii * ny + jj) * nz + kk;
                    const int64_t d = lay.offset(q, n);
                    cols[d] = (DstI)c;
                    vals[d] = (di == 0 && dj == 0 && dk == 0) ? cval : oval;
                    lo = c < lo ? c : lo; hi = c > hi ? c : hi;
                    n++;
                }
        const int64_t pad = r < ncols ? r : ncols - 1;
        for (; n < 27; n++) {
            const int64_t d = lay.offset(q, n);
            cols[d] = (DstI)pad; vals[d] = 0.0;
            lo = pad < lo ? pad : lo; hi = pad > hi ? pad : hi;
        }
    }
    if (minmax) block_minmax(lo, hi, minmax);
}

template <typename DstI>
__global__ void gen_random_kernel(int64_t num_columns, int K, uint64_t seed,
                                  DstI *__restrict__ cols, double *__restrict__ vals,
                                  EllLayout lay, int64_t row_begin, long long *minmax)
{
    long long lo = 0x7fffffffffffffffLL, hi = -1;
    for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < lay.num_rows;
         q += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = row_begin + q;
        for (int l = 0; l < K; l++) {
            const uint64_t u = splitmix64(seed ^ (uint64_t)(r * K + l));
            const int64_t c = (int64_t)__umul64hi(u, (uint64_t)num_columns);
            const double v = __dadd_rn(__dmul_rn(2.0, __dmul_rn((double)(splitmix64(u) >> 11), 0x1.0p-53)), -1.0);
            const int64_t d = lay.offset(q, l);
            cols[d] = (DstI)c; vals[d] = v;
            lo = c < lo ? c : lo; hi = c > hi ? c : hi;
        }
    }
    if (minmax) block_minmax(lo, hi, minmax);
}

template <typename DstI>
static cudaError_t generate_typed(int kind, const int64_t dims[3], const double vals[2], uint64_t seed,
                                  DstI *dst_cols, double *dst_vals, const EllLayout &lay,
                                  int64_t row_begin, long long *minmax, cudaStream_t stream)
{
    const int g = grid_for(lay.num_rows, 128);
    switch (kind) {
    case ELLSPMV_CUDA_GEN_LAPLACE2D:
        gen_laplace2d_kernel<DstI><<<g, 128, 0, stream>>>(dims[0], dims[1], vals[0], vals[1], dst_cols, dst_vals, lay, row_begin, minmax);
        break;
    case ELLSPMV_CUDA_GEN_STENCIL27:
        gen_stencil27_kernel<DstI><<<g, 128, 0, stream>>>(dims[0], dims[1], dims[2], vals[0], vals[1], dst_cols, dst_vals, lay, row_begin, minmax);
        break;
    case ELLSPMV_CUDA_GEN_RANDOM:
        gen_random_kernel<DstI><<<g, 128, 0, stream>>>(dims[1], (int)dims[2], seed, dst_cols, dst_vals, lay, row_begin, minmax);
        break;
    default:
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

cudaError_t generate_sliced(int kind, const int64_t dims[3], const double vals[2], uint64_t seed,
                            int dst_idx_bits, void *dst_cols, double *dst_vals,
                            const EllLayout &lay, int64_t row_begin,
                            long long *minmax, cudaStream_t stream)
{
    if (lay.num_rows <= 0) return cudaSuccess;
    if (dst_idx_bits == 32)
        return generate_typed<int32_t>(kind, dims, vals, seed, (int32_t *)dst_cols, dst_vals, lay, row_begin, minmax, stream);
    return generate_typed<int64_t>(kind, dims, vals, seed, (int64_t *)dst_cols, dst_vals, lay, row_begin, minmax, stream);
}

// CSR view of the random matrix: every row has exactly K entries, so
// rowptr[i] = i*K and the entry arrays are the row-major ELL arrays.
template <typename DstI>
__global__ void gen_csr_random_kernel(int64_t num_rows, int64_t num_columns, int K, uint64_t seed,
                                      int64_t *__restrict__ rowptr, DstI *__restrict__ cols,
                                      double *__restrict__ vals)
{
    const int64_t n = num_rows * K;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n;
         e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t u = splitmix64(seed ^ (uint64_t)e);
        cols[e] = (DstI)__umul64hi(u, (uint64_t)num_columns);
        vals[e] = __dadd_rn(__dmul_rn(2.0, __dmul_rn((double)(splitmix64(u) >> 11), 0x1.0p-53)), -1.0);
    }
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= num_rows;
         r += (int64_t)gridDim.x * blockDim.x)
        rowptr[r] = r * K;
}

cudaError_t generate_csr_random(const int64_t dims[3], uint64_t seed, int idx_bits,
                                int64_t *rowptr, void *cols, double *vals, cudaStream_t stream)
{
    const int64_t n = dims[0] * dims[2];
    const int g = grid_for(n > dims[0] + 1 ? n : dims[0] + 1, 256);
    if (idx_bits == 32)
        gen_csr_random_kernel<int32_t><<<g, 256, 0, stream>>>(dims[0], dims[1], (int)dims[2], seed, rowptr, (int32_t *)cols, vals);
    else
        gen_csr_random_kernel<int64_t><<<g, 256, 0, stream>>>(dims[0], dims[1], (int)dims[2], seed, rowptr, (int64_t *)cols, vals);
    return cudaGetLastError();
}

// CSR form of the two stencils: what csr_from_coo (csrspmv.c:1390-1475) makes of the same
// canonical COO stream -- the entries of a row in the ELL order above, no padding.  Row lengths
// are products (27-point) or sums (5-point) of per-axis neighbour counts, so rowptr has a closed
// form and one thread per row writes its pointer and its entries without a scan.
//   axis_prefix(t, n) = sum over t' < t of (1 + [t' > 0] + [t' < n-1])
__host__ __device__ __forceinline__ int64_t axis_prefix(int64_t t, int64_t n)
{
    return t + (t > 0 ? t - 1 : 0) + (t < n - 1 ? t : n - 1);
}
__host__ __device__ __forceinline__ int64_t laplace2d_rowptr(int64_t i, int64_t j, int64_t nx, int64_t ny)
{
    const int64_t ci = (i > 0) + (i + 1 < nx);
    const int64_t cj_before = (j > 0 ? j - 1 : 0) + (j < ny - 1 ? j : ny - 1);
    return ny * axis_prefix(i, nx) + i * (2 * ny - 2) + j * (1 + ci) + cj_before;
}
__host__ __device__ __forceinline__ int64_t stencil27_rowptr(int64_t i, int64_t j, int64_t k, int64_t nx, int64_t ny, int64_t nz)
{
    const int64_t a = 1 + (i > 0) + (i + 1 < nx), b = 1 + (j > 0) + (j + 1 < ny);
    return axis_prefix(i, nx) * (3 * ny - 2) * (3 * nz - 2) + a * axis_prefix(j, ny) * (3 * nz - 2) + a * b * axis_prefix(k, nz);
}

int64_t csr_stencil_nnz(int kind, const int64_t dims[3])
{
    if (kind == ELLSPMV_CUDA_GEN_LAPLACE2D)
        return dims[0] > 0 && dims[1] > 0 ? laplace2d_rowptr(dims[0], 0, dims[0], dims[1]) : 0;
    if (kind == ELLSPMV_CUDA_GEN_STENCIL27)
        return dims[0] > 0 && dims[1] > 0 && dims[2] > 0 ? (3 * dims[0] - 2) * (3 * dims[1] - 2) * (3 * dims[2] - 2) : 0;
    return -1;
}

template <typename DstI>
__global__ void gen_csr_laplace2d_kernel(int64_t nx, int64_t ny, double cval, double oval, int64_t *__restrict__ rowptr,
                                         DstI *__restrict__ cols, double *__restrict__ vals)
{
    const int64_t rows = nx * ny;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = r / ny, j = r - i * ny;
        int64_t e = laplace2d_rowptr(i, j, nx, ny);
        rowptr[r] = e;
        if (r == rows) break;
        if (i > 0)      { cols[e] = (DstI)(r - ny); vals[e] = oval; e++; }
        if (j > 0)      { cols[e] = (DstI)(r - 1);  vals[e] = oval; e++; }
        cols[e] = (DstI)r; vals[e] = cval; e++;
        if (j + 1 < ny) { cols[e] = (DstI)(r + 1);  vals[e] = oval; e++; }
        if (i + 1 < nx) { cols[e] = (DstI)(r + ny); vals[e] = oval; e++; }
    }
}

template <typename DstI>
__global__ void gen_csr_stencil27_kernel(int64_t nx, int64_t ny, int64_t nz, double cval, double oval,
                                         int64_t *__restrict__ rowptr, DstI *__restrict__ cols, double *__restrict__ vals)
{
    const int64_t rows = nx * ny * nz;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r <= rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = r % nz, j = (r / nz) % ny, i = r / (nz * ny);
        int64_t e = stencil27_rowptr(i, j, k, nx, ny, nz);
        rowptr[r] = e;
        if (r == rows) break;
        for (int di = -1; di <= 1; di++)
            for (int dj = -1; dj <= 1; dj++)
                for (int dk = -1; dk <= 1; dk++) {
                    const int64_t ii = i + di, jj = j + dj, kk = k + dk;
                    if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
                    cols[e] = (DstI)((ii * ny + jj) * nz + kk);
                    vals[e] = (di == 0 && dj == 0 && dk == 0) ? cval : oval;
                    e++;
                }
    }
}

cudaError_t generate_csr_stencil(int kind, const int64_t dims[3], const double v[2], int idx_bits,
                                 int64_t *rowptr, void *cols, double *vals, cudaStream_t stream)
{
    const int64_t rows = kind == ELLSPMV_CUDA_GEN_LAPLACE2D ? dims[0] * dims[1] : dims[0] * dims[1] * dims[2];
    const int g = grid_for(rows + 1, 128);
    if (kind == ELLSPMV_CUDA_GEN_LAPLACE2D) {
        if (idx_bits == 32) gen_csr_laplace2d_kernel<int32_t><<<g, 128, 0, stream>>>(dims[0], dims[1], v[0], v[1], rowptr, (int32_t *)cols, vals);
        else gen_csr_laplace2d_kernel<int64_t><<<g, 128, 0, stream>>>(dims[0], dims[1], v[0], v[1], rowptr, (int64_t *)cols, vals);
    } else if (kind == ELLSPMV_CUDA_GEN_STENCIL27) {
        if (idx_bits == 32) gen_csr_stencil27_kernel<int32_t><<<g, 128, 0, stream>>>(dims[0], dims[1], dims[2], v[0], v[1], rowptr, (int32_t *)cols, vals);
        else gen_csr_stencil27_kernel<int64_t><<<g, 128, 0, stream>>>(dims[0], dims[1], dims[2], v[0], v[1], rowptr, (int64_t *)cols, vals);
    } else {
        return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

// ---- which slices read columns outside the shard's own rows (fused step sync, ell_kernels.cu) ----
// remote[s] = 1 when slice s references a column outside [lo, hi): such a CTA must wait for the
// peers' pushes of the previous step before it gathers.  lo/hi are the shard's row range shrunk
// to multiples of 16 entries (one 128-byte line of x), so a CTA that does not wait never pulls a
// line into L1 that also holds entries a peer is still writing.
template <typename IdxT>
__global__ void __launch_bounds__(kBlockThreads)
slice_remote_kernel(const IdxT *__restrict__ cols, EllLayout lay, long long lo, long long hi,
                    unsigned char *__restrict__ remote)
{
    const int64_t s = blockIdx.x;
    const int n = lay.slice_rows * lay.rowsize;
    const IdxT *c = cols + s * (int64_t)n;
    int any = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long v = (long long)c[i];
        any |= (v < lo || v >= hi);
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) remote[s] = any ? 1 : 0;
}

cudaError_t mark_remote_slices(int idx_bits, const void *cols, const EllLayout &lay, int64_t lo, int64_t hi,
                               unsigned char *remote, cudaStream_t stream)
{
    if (lay.num_slices <= 0) return cudaSuccess;
    if (lay.num_slices > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (idx_bits == 64)
        slice_remote_kernel<int64_t><<<(unsigned)lay.num_slices, kBlockThreads, 0, stream>>>((const int64_t *)cols, lay, lo, hi, remote);
    else
        slice_remote_kernel<int32_t><<<(unsigned)lay.num_slices, kBlockThreads, 0, stream>>>((const int32_t *)cols, lay, lo, hi, remote);
    return cudaGetLastError();
}

// ---- CSR rows -> sliced ELL (the ELL view of a balanced CSR matrix, api.cu) --------------------
// One CTA per slice: its rows' entries are one contiguous run of the CSR arrays, read coalesced;
// each entry finds its row by bisection in the slice's row pointers (shared memory).
template <typename SrcI, typename DstI>
__global__ void __launch_bounds__(kBlockThreads)
csr_to_sliced_kernel(const int64_t *__restrict__ rowptr, const SrcI *__restrict__ src_cols,
                     const double *__restrict__ src_vals, DstI *__restrict__ dst_cols,
                     double *__restrict__ dst_vals, int *__restrict__ rowlen, EllLayout lay)
{
    extern __shared__ long long s_rp[];              // slice_rows + 1 row pointers
    const int S = lay.slice_rows, K = lay.rowsize;
    const int64_t s = blockIdx.x;
    const int64_t r0 = s * S;
    const int nr = (int)((r0 + S <= lay.num_rows) ? S : lay.num_rows - r0);
    for (int i = threadIdx.x; i <= nr; i += blockDim.x) s_rp[i] = rowptr[r0 + i];
    __syncthreads();
    const int64_t kb = s_rp[0], ke = s_rp[nr];
    const int64_t base = s * S * (int64_t)K;
    for (int64_t k = kb + threadIdx.x; k < ke; k += blockDim.x) {
        int lo = 0, hi = nr;                          // last row with s_rp[row] <= k
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_rp[mid] <= k) lo = mid; else hi = mid;
        }
        const int l = (int)(k - s_rp[lo]);
        const int64_t d = base + (int64_t)l * S + lo;
        dst_cols[d] = (DstI)src_cols[k];
        dst_vals[d] = src_vals[k];
    }
    for (int r = threadIdx.x; r < nr; r += blockDim.x) {
        const int len = (int)(s_rp[r + 1] - s_rp[r]);
        if (rowlen) rowlen[r0 + r] = len;
        const DstI pad = len > 0 ? (DstI)src_cols[s_rp[r + 1] - 1] : (DstI)0;
        for (int l = len; l < K; l++) {
            const int64_t d = base + (int64_t)l * S + r;
            dst_cols[d] = pad;
            dst_vals[d] = 0.0;
        }
    }
}

cudaError_t csr_to_sliced(int src_idx_bits, int dst_idx_bits, const int64_t *rowptr, const void *src_cols,
                          const double *src_vals, void *dst_cols, double *dst_vals, int *rowlen,
                          const EllLayout &lay, cudaStream_t stream)
{
    if (lay.num_slices <= 0) return cudaSuccess;
    if (lay.num_slices > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned g = (unsigned)lay.num_slices;
    const size_t smem = (size_t)(lay.slice_rows + 1) * 8;
    if (src_idx_bits == 32 && dst_idx_bits == 32)
        csr_to_sliced_kernel<int32_t, int32_t><<<g, kBlockThreads, smem, stream>>>(rowptr, (const int32_t *)src_cols, src_vals, (int32_t *)dst_cols, dst_vals, rowlen, lay);
    else if (src_idx_bits == 64 && dst_idx_bits == 64)
        csr_to_sliced_kernel<int64_t, int64_t><<<g, kBlockThreads, smem, stream>>>(rowptr, (const int64_t *)src_cols, src_vals, (int64_t *)dst_cols, dst_vals, rowlen, lay);
    else if (src_idx_bits == 64 && dst_idx_bits == 32)
        csr_to_sliced_kernel<int64_t, int32_t><<<g, kBlockThreads, smem, stream>>>(rowptr, (const int64_t *)src_cols, src_vals, (int32_t *)dst_cols, dst_vals, rowlen, lay);
    else
        return cudaErrorInvalidValue;
    return cudaGetLastError();
}

// ---- largest column each row chunk of the pipelined host call references (api.cu) -------------
template <typename IdxT>
__global__ void __launch_bounds__(kBlockThreads)
chunk_max_kernel(const IdxT *__restrict__ cols, EllLayout lay, int64_t chunk_slices, long long *__restrict__ out)
{
    const int64_t s = blockIdx.x;
    const int n = lay.slice_rows * lay.rowsize;
    const IdxT *c = cols + s * (int64_t)n;
    long long best = -1;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const long long v = (long long)c[i];
        best = v > best ? v : best;
    }
    for (int off = 16; off > 0; off >>= 1) {
        const long long o = __shfl_xor_sync(0xffffffffu, best, off);
        best = o > best ? o : best;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(out + s / chunk_slices, best);
}

cudaError_t chunk_max_cols(int idx_bits, const void *cols, const EllLayout &lay, int64_t chunk_slices, int nchunks,
                           long long *d_out, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_out, 0xff, (size_t)nchunks * 8, stream);      // -1
    if (e != cudaSuccess || lay.num_slices <= 0) return e;
    if (idx_bits == 64)
        chunk_max_kernel<int64_t><<<(unsigned)lay.num_slices, kBlockThreads, 0, stream>>>((const int64_t *)cols, lay, chunk_slices, d_out);
    else
        chunk_max_kernel<int32_t><<<(unsigned)lay.num_slices, kBlockThreads, 0, stream>>>((const int32_t *)cols, lay, chunk_slices, d_out);
    return cudaGetLastError();
}

}  // namespace ellspmv
