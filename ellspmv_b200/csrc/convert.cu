// convert.cu -- COO -> sliced ELL / CSR on the device.
//
// Same result, bit for bit, as the reference's serial converters followed by
// upload: ell_from_coo_size / ell_from_coo (ellspmv.c:931-958, 1081-1127) and
// csr_from_coo (csrspmv.c:1436-1465, general branch).  Those place entry k of
// the file in the next free slot of its row, i.e. they are a STABLE sort of
// the entries by row.  Here: a stable LSD radix sort of (row, file position)
// pairs (cub::DeviceRadixSort -- library code, this is conversion, not the hot
// path), a histogram + scan for the row starts, then one scatter kernel that
// writes slot = rank-in-row straight into the device layout and one kernel
// for the reference's padding rule (column min(i, ncols-1), value 0.0).
#include <cub/cub.cuh>

#include "common.cuh"

namespace ellspmv {

template <typename IdxT>
__global__ void coo_keys_kernel(const IdxT *__restrict__ rowidx, int64_t nnz, int64_t num_rows,
                                IdxT *__restrict__ keys, int64_t *__restrict__ pos,
                                unsigned long long *__restrict__ counts, int *__restrict__ bad)
{
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < nnz; k += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = (int64_t)rowidx[k] - 1;            // 1-based in the file
        if (r < 0 || r >= num_rows) { *bad = 1; keys[k] = 0; pos[k] = k; continue; }
        keys[k] = (IdxT)r;
        pos[k] = k;
        atomicAdd(&counts[r], 1ULL);
    }
}

template <typename IdxT, typename DstI>
__global__ void coo_scatter_ell_kernel(const IdxT *__restrict__ keys_sorted, const int64_t *__restrict__ perm,
                                       const int64_t *__restrict__ rowstart, const IdxT *__restrict__ colidx,
                                       const double *__restrict__ a, int64_t nnz, int64_t num_columns,
                                       DstI *__restrict__ dst_cols, double *__restrict__ dst_vals, EllLayout lay,
                                       long long *minmax, int *__restrict__ bad)
{
    long long lo = 0x7fffffffffffffffLL, hi = -1;
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = keys_sorted[p];
        const int64_t k = perm[p];
        const int slot = (int)(p - rowstart[r]);
        const long long c = (long long)colidx[k] - 1;
        if (c < 0 || c >= num_columns) { *bad = 1; continue; }
        const int64_t d = lay.offset(r, slot);
        dst_cols[d] = (DstI)c;
        dst_vals[d] = a[k];
        lo = c < lo ? c : lo; hi = c > hi ? c : hi;
    }
    // block_minmax (layout.cu) inlined: one atomic pair per warp
    for (int off = 16; off > 0; off >>= 1) {
        long long olo = __shfl_xor_sync(0xffffffffu, lo, off), ohi = __shfl_xor_sync(0xffffffffu, hi, off);
        lo = olo < lo ? olo : lo; hi = ohi > hi ? ohi : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
}

template <typename DstI>
__global__ void ell_pad_kernel(const unsigned long long *__restrict__ counts, int64_t num_columns,
                               DstI *__restrict__ dst_cols, double *__restrict__ dst_vals, EllLayout lay,
                               long long *minmax)
{
    long long lo = 0x7fffffffffffffffLL, hi = -1;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < lay.num_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const long long pad = r < num_columns ? r : num_columns - 1;
        for (int l = (int)counts[r]; l < lay.rowsize; l++) {
            const int64_t d = lay.offset(r, l);
            dst_cols[d] = (DstI)pad;
            dst_vals[d] = 0.0;
            lo = pad < lo ? pad : lo; hi = pad > hi ? pad : hi;
        }
    }
    for (int off = 16; off > 0; off >>= 1) {
        long long olo = __shfl_xor_sync(0xffffffffu, lo, off), ohi = __shfl_xor_sync(0xffffffffu, hi, off);
        lo = olo < lo ? olo : lo; hi = ohi > hi ? ohi : hi;
    }
    if ((threadIdx.x & 31) == 0 && lo <= hi) { atomicMin(minmax, lo); atomicMax(minmax + 1, hi); }
}

template <typename IdxT>
__global__ void coo_gather_csr_kernel(const int64_t *__restrict__ perm, const IdxT *__restrict__ colidx,
                                      const double *__restrict__ a, int64_t nnz, int64_t num_columns,
                                      IdxT *__restrict__ csrcolidx, double *__restrict__ csra, int *__restrict__ bad)
{
    for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < nnz; p += (int64_t)gridDim.x * blockDim.x) {
        const int64_t k = perm[p];
        const long long c = (long long)colidx[k] - 1;
        if (c < 0 || c >= num_columns) { *bad = 1; csrcolidx[p] = 0; csra[p] = 0.0; continue; }
        csrcolidx[p] = (IdxT)c;
        csra[p] = a[k];
    }
}

static int grid_of(int64_t n) { int64_t g = (n + 255) / 256; return (int)(g > 148 * 32 ? 148 * 32 : (g < 1 ? 1 : g)); }

// Device-side state shared by the ELL and CSR routes.
template <typename IdxT>
struct CooSorted {
    IdxT *keys = nullptr, *keys_sorted = nullptr;
    int64_t *pos = nullptr, *perm = nullptr, *rowstart = nullptr;   // rowstart: num_rows + 1
    unsigned long long *counts = nullptr;
    int *bad = nullptr;
    void *temp = nullptr;
    unsigned long long maxcount = 0;
    void release() {
        cudaFree(keys); cudaFree(keys_sorted); cudaFree(pos); cudaFree(perm); cudaFree(rowstart);
        cudaFree(counts); cudaFree(bad); cudaFree(temp);
    }
};

template <typename IdxT>
static cudaError_t sort_coo(CooSorted<IdxT> &w, const IdxT *d_rowidx, int64_t nnz, int64_t num_rows,
                            cudaStream_t stream, int *host_bad)
{
    cudaError_t e;
    const size_t nz = (size_t)(nnz > 0 ? nnz : 1), nr = (size_t)num_rows + 1;
    if ((e = cudaMalloc(&w.keys, nz * sizeof(IdxT))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.keys_sorted, nz * sizeof(IdxT))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.pos, nz * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.perm, nz * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.rowstart, nr * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.counts, nr * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&w.bad, sizeof(int))) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w.counts, 0, nr * 8, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(w.bad, 0, sizeof(int), stream)) != cudaSuccess) return e;
    if (nnz > 0) coo_keys_kernel<IdxT><<<grid_of(nnz), 256, 0, stream>>>(d_rowidx, nnz, num_rows, w.keys, w.pos, w.counts, w.bad);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    // stable sort of (row, file position) by row; only the bits a row index can use
    int bits = 1;
    while (bits < (int)sizeof(IdxT) * 8 - 1 && (((int64_t)1 << bits) < num_rows)) bits++;
    size_t temp_bytes = 0, tb2 = 0, tb3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, w.keys, w.keys_sorted, w.pos, w.perm, (int64_t)nnz, 0, bits, stream);
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, w.counts, (unsigned long long *)w.rowstart, (int64_t)nr, stream);
    unsigned long long *d_max = nullptr;
    cub::DeviceReduce::Max(nullptr, tb3, w.counts, d_max, (int64_t)nr, stream);
    if (tb2 > temp_bytes) temp_bytes = tb2;
    if (tb3 > temp_bytes) temp_bytes = tb3;
    if ((e = cudaMalloc(&w.temp, temp_bytes + 16)) != cudaSuccess) return e;
    if (nnz > 0) {
        e = cub::DeviceRadixSort::SortPairs(w.temp, temp_bytes, w.keys, w.keys_sorted, w.pos, w.perm, (int64_t)nnz, 0, bits, stream);
        if (e != cudaSuccess) return e;
    }
    e = cub::DeviceScan::ExclusiveSum(w.temp, temp_bytes, w.counts, (unsigned long long *)w.rowstart, (int64_t)nr, stream);
    if (e != cudaSuccess) return e;
    // the maximum goes into the spare slot counts[num_rows] is NOT safe (it is an input); use pos[0..] instead
    if ((e = cudaMalloc(&d_max, 8)) != cudaSuccess) return e;
    e = cub::DeviceReduce::Max(w.temp, temp_bytes, w.counts, d_max, (int64_t)nr, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&w.maxcount, d_max, 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(host_bad, w.bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d_max);
    return e;
}

// ---- entry points used by api.cu (CooEllJob is declared in common.cuh) -----------------
// Phase 1: sort + K.  Phase 2 (after the caller allocated the sliced arrays): scatter + pad.
template <typename IdxT>
static cudaError_t ell_phase1(CooEllJob &job, cudaStream_t stream, int *bad)
{
    auto *w = new CooSorted<IdxT>();
    job.state = w;
    cudaError_t e = sort_coo<IdxT>(*w, (const IdxT *)job.d_rowidx, job.nnz, job.num_rows, stream, bad);
    job.rowsize = (int64_t)w->maxcount;
    return e;
}

template <typename IdxT, typename DstI>
static cudaError_t ell_phase2(CooEllJob &job, DstI *dst_cols, double *dst_vals, const EllLayout &lay,
                              long long *minmax, cudaStream_t stream, int *bad)
{
    auto *w = (CooSorted<IdxT> *)job.state;
    if (job.nnz > 0)
        coo_scatter_ell_kernel<IdxT, DstI><<<grid_of(job.nnz), 256, 0, stream>>>(
            w->keys_sorted, w->perm, w->rowstart, (const IdxT *)job.d_colidx, job.d_a, job.nnz, job.num_columns,
            dst_cols, dst_vals, lay, minmax, w->bad);
    if (lay.num_rows > 0 && lay.rowsize > 0)
        ell_pad_kernel<DstI><<<grid_of(lay.num_rows), 256, 0, stream>>>(w->counts, job.num_columns, dst_cols, dst_vals, lay, minmax);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(bad, w->bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    return e;
}

cudaError_t coo_to_ell_phase1(CooEllJob &job, cudaStream_t stream, int *bad)
{
    return job.idx_bits == 64 ? ell_phase1<int64_t>(job, stream, bad) : ell_phase1<int32_t>(job, stream, bad);
}

cudaError_t coo_to_ell_phase2(CooEllJob &job, int dst_idx_bits, void *dst_cols, double *dst_vals,
                              const EllLayout &lay, long long *minmax, cudaStream_t stream, int *bad)
{
    if (job.idx_bits == 64) {
        if (dst_idx_bits == 64) return ell_phase2<int64_t, int64_t>(job, (int64_t *)dst_cols, dst_vals, lay, minmax, stream, bad);
        return ell_phase2<int64_t, int32_t>(job, (int32_t *)dst_cols, dst_vals, lay, minmax, stream, bad);
    }
    return ell_phase2<int32_t, int32_t>(job, (int32_t *)dst_cols, dst_vals, lay, minmax, stream, bad);
}

void coo_to_ell_release(CooEllJob &job)
{
    if (!job.state) return;
    if (job.idx_bits == 64) { auto *w = (CooSorted<int64_t> *)job.state; w->release(); delete w; }
    else { auto *w = (CooSorted<int32_t> *)job.state; w->release(); delete w; }
    job.state = nullptr;
}

// CSR: rowptr (num_rows+1 int64), colidx, a written into caller-allocated device arrays
template <typename IdxT>
static cudaError_t csr_all(const IdxT *d_rowidx, const IdxT *d_colidx, const double *d_a, int64_t nnz,
                           int64_t num_rows, int64_t num_columns, int64_t *rowptr, IdxT *csrcolidx, double *csra,
                           cudaStream_t stream, int *bad)
{
    CooSorted<IdxT> w;
    cudaError_t e = sort_coo<IdxT>(w, d_rowidx, nnz, num_rows, stream, bad);
    if (e == cudaSuccess) e = cudaMemcpyAsync(rowptr, w.rowstart, (size_t)(num_rows + 1) * 8, cudaMemcpyDeviceToDevice, stream);
    if (e == cudaSuccess && nnz > 0) {
        coo_gather_csr_kernel<IdxT><<<grid_of(nnz), 256, 0, stream>>>(w.perm, d_colidx, d_a, nnz, num_columns, csrcolidx, csra, w.bad);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(bad, w.bad, sizeof(int), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    w.release();
    return e;
}

cudaError_t coo_to_csr(int idx_bits, const void *d_rowidx, const void *d_colidx, const double *d_a, int64_t nnz,
                       int64_t num_rows, int64_t num_columns, int64_t *rowptr, void *csrcolidx, double *csra,
                       cudaStream_t stream, int *bad)
{
    if (idx_bits == 64)
        return csr_all<int64_t>((const int64_t *)d_rowidx, (const int64_t *)d_colidx, d_a, nnz, num_rows, num_columns,
                                rowptr, (int64_t *)csrcolidx, csra, stream, bad);
    return csr_all<int32_t>((const int32_t *)d_rowidx, (const int32_t *)d_colidx, d_a, nnz, num_rows, num_columns,
                            rowptr, (int32_t *)csrcolidx, csra, stream, bad);
}

}  // namespace ellspmv
