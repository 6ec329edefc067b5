// ell_blocked.cu -- column-blocked ELL for matrices whose x does not fit in L2
// (ELLSPMV_CUDA_COLUMN_BLOCKED; tolerance mode).
//
// Why: on B200 a random 8-byte gather that misses L2 costs a ~100-byte line fill
// from HBM and tops out at 67 G gathers/s (profiles/r1_c4_gather.md), so BASELINE
// config 4 (random 50M x 32, x = 400 MB) runs 9x above its algorithmic bytes.
// Cache blocking fixes the locality instead of the kernel: the columns are cut
// into B blocks whose slice of x (<= 48 MB) stays resident in the 126 MB L2, the
// entries are binned by block, and y += A_b * x runs block after block.
//
// Layout: rows are grouped in micro-slices of 32 (one warp).  For block b and
// micro-slice m, the 32 rows are ordered by their number of entries in block b
// (descending, stable) and stored as jagged diagonals: "diagonal" j holds the
// j-th entry of every row that has more than j entries, in that order, so the
// rows still active at step j are exactly lanes 0..n_j-1 and a warp's load is
// one contiguous run -- coalesced with NO padding at all (a plain per-block ELL
// pads a uniform random matrix by 1.8x).  Per (b, m): a 64-bit start offset and
// 32 x 16-bit (count, original lane) records.  Stored zeros (the ELL padding)
// are dropped: 0*x contributes nothing for finite x.
//
// Arithmetic: inside a block a row's entries are added in the reference's
// order; the per-block partial sums are then added block by block.  That is a
// different association than the reference's single left-to-right chain, hence
// tolerance mode (bound in tests/test_gpu_ell.py).  One launch per block, in
// stream order, keeps the result deterministic; each launch carries an L2
// persisting access-policy window over its slice of x so that the streaming
// matrix cannot evict it.
#include <cub/cub.cuh>

#include "common.cuh"

namespace ellspmv {

constexpr int kMaxBlocks = 64;

struct CbLayout {
    int num_blocks = 0;
    int64_t block_cols = 0;      // W: columns per block
    int64_t num_micro = 0;       // ceil(rows / 32)
    int64_t total_entries = 0;   // padded entries over all blocks
};

// ---- build ---------------------------------------------------------------------
// nonzero entries of `row` in column block b, and this lane's position when the 32
// rows are sorted by that count (descending, ties by lane)
__device__ __forceinline__ int sorted_position(int cnt, int lane)
{
    int pos = 0;
    for (int i = 0; i < 32; i++) {
        const int ci = __shfl_sync(0xffffffffu, cnt, i);
        pos += (ci > cnt) || (ci == cnt && i < lane);
    }
    return pos;
}

// pass 1: per (block, micro-slice) the entry count and the 32 (count, lane) records
template <typename IdxT>
__global__ void cb_count_kernel(const double *__restrict__ vals, const IdxT *__restrict__ cols, EllLayout lay,
                                int num_blocks, int64_t block_cols, int64_t num_micro,
                                long long *__restrict__ sizes, unsigned short *__restrict__ meta)
{
    const int lane = threadIdx.x & 31;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= num_micro) return;
    const int64_t row = m * 32 + lane;
    unsigned char cnt[kMaxBlocks];
    for (int b = 0; b < num_blocks; b++) cnt[b] = 0;
    if (row < lay.num_rows) {
        for (int l = 0; l < lay.rowsize; l++) {
            const int64_t o = lay.offset(row, l);
            if (vals[o] != 0.0) cnt[(int)((int64_t)cols[o] / block_cols)]++;
        }
    }
    for (int b = 0; b < num_blocks; b++) {
        const int c = cnt[b];
        const int pos = sorted_position(c, lane);
        const int64_t i = (int64_t)b * num_micro + m;
        meta[i * 32 + pos] = (unsigned short)(c | (lane << 8));
        int total = c;
        for (int off = 16; off > 0; off >>= 1) total += __shfl_xor_sync(0xffffffffu, total, off);
        if (lane == 0) sizes[i] = total;
    }
}

// pass 2: write the jagged diagonals
template <typename IdxT>
__global__ void cb_fill_kernel(const double *__restrict__ vals, const IdxT *__restrict__ cols, EllLayout lay,
                               int num_blocks, int64_t block_cols, int64_t num_micro,
                               const long long *__restrict__ offs, double *__restrict__ cb_vals, IdxT *__restrict__ cb_cols)
{
    const int lane = threadIdx.x & 31;
    const int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= num_micro) return;
    const int64_t row = m * 32 + lane;
    const bool live = row < lay.num_rows;
    for (int b = 0; b < num_blocks; b++) {
        int cnt = 0;
        if (live)
            for (int l = 0; l < lay.rowsize; l++) {
                const int64_t o = lay.offset(row, l);
                cnt += (vals[o] != 0.0) && ((int)((int64_t)cols[o] / block_cols) == b);
            }
        const int pos = sorted_position(cnt, lane);
        int maxc = cnt;
        for (int off = 16; off > 0; off >>= 1) { int o = __shfl_xor_sync(0xffffffffu, maxc, off); maxc = o > maxc ? o : maxc; }
        long long base = offs[(int64_t)b * num_micro + m];
        int cursor = 0;
        for (int j = 0; j < maxc; j++) {
            const unsigned active = __ballot_sync(0xffffffffu, cnt > j);
            if (cnt > j) {
                // this row's next entry that belongs to block b
                for (;; cursor++) {
                    const int64_t o = lay.offset(row, cursor);
                    const double v = vals[o];
                    const IdxT c = cols[o];
                    if (v != 0.0 && (int)((int64_t)c / block_cols) == b) {
                        cb_vals[base + pos] = v;
                        cb_cols[base + pos] = c;
                        cursor++;
                        break;
                    }
                }
            }
            base += __popc(active);
        }
    }
}

struct CbMatrix {
    CbLayout lay;
    int idx_bits = 32;
    double *vals = nullptr;
    void *cols = nullptr;
    long long *offs = nullptr;        // num_blocks * num_micro + 1
    unsigned short *meta = nullptr;   // num_blocks * num_micro * 32: count | lane << 8, sorted by count
    int64_t bytes = 0;
    int num_sms = 148;
    bool persist = false;             // an L2 persisting carve-out for the x block is configured
};

void cb_free(CbMatrix *cb)
{
    if (!cb) return;
    cudaFree(cb->vals); cudaFree(cb->cols); cudaFree(cb->offs); cudaFree(cb->meta);
    delete cb;
}

template <typename IdxT>
static cudaError_t cb_build_typed(CbMatrix *cb, const double *vals, const IdxT *cols, const EllLayout &lay,
                                  cudaStream_t stream)
{
    const CbLayout &L = cb->lay;
    const int64_t n = (int64_t)L.num_blocks * L.num_micro;
    cudaError_t e;
    long long *sizes = nullptr;
    void *temp = nullptr;
    if ((e = cudaMalloc(&sizes, (size_t)(n + 1) * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&cb->offs, (size_t)(n + 1) * 8)) != cudaSuccess) { cudaFree(sizes); return e; }
    if ((e = cudaMalloc(&cb->meta, (size_t)n * 32 * sizeof(unsigned short))) != cudaSuccess) { cudaFree(sizes); return e; }
    e = cudaMemsetAsync(sizes, 0, (size_t)(n + 1) * 8, stream);
    const int threads = 128;
    const int64_t grid = (L.num_micro * 32 + threads - 1) / threads;
    if (e == cudaSuccess && grid > 0) {
        cb_count_kernel<IdxT><<<(unsigned)grid, threads, 0, stream>>>(vals, cols, lay, L.num_blocks, L.block_cols, L.num_micro, sizes, cb->meta);
        e = cudaGetLastError();
    }
    size_t temp_bytes = 0;
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, sizes, cb->offs, n + 1, stream);
    if (e == cudaSuccess) e = cudaMalloc(&temp, temp_bytes + 16);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, sizes, cb->offs, n + 1, stream);
    long long total = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&total, cb->offs + n, 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(sizes);
    cudaFree(temp);
    if (e != cudaSuccess) return e;
    cb->lay.total_entries = total;
    const size_t ne = (size_t)(total > 0 ? total : 1);
    if ((e = cudaMalloc(&cb->vals, ne * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&cb->cols, ne * sizeof(IdxT))) != cudaSuccess) return e;
    cb->bytes = (int64_t)(ne * (8 + sizeof(IdxT)) + (size_t)(n + 1) * 8 + (size_t)n * 64);
    if (grid > 0) {
        cb_fill_kernel<IdxT><<<(unsigned)grid, threads, 0, stream>>>(vals, cols, lay, L.num_blocks, L.block_cols, L.num_micro,
                                                                    cb->offs, cb->vals, (IdxT *)cb->cols);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    return e;
}

// Build from the regular sliced-ELL arrays.  *out = nullptr (and success) when
// blocking does not apply (x already fits the target).
cudaError_t cb_build(CbMatrix **out, int idx_bits, const double *vals, const void *cols, const EllLayout &lay,
                     int64_t num_columns, int64_t target_x_bytes, cudaStream_t stream)
{
    *out = nullptr;
    if (lay.num_rows <= 0 || lay.rowsize <= 0 || lay.rowsize > 255) return cudaSuccess;
    int64_t nb = (num_columns * 8 + target_x_bytes - 1) / target_x_bytes;
    if (nb <= 1) return cudaSuccess;
    if (nb > kMaxBlocks) nb = kMaxBlocks;
    CbMatrix *cb = new (std::nothrow) CbMatrix();
    if (!cb) return cudaErrorMemoryAllocation;
    cb->idx_bits = idx_bits;
    cb->lay.num_blocks = (int)nb;
    cb->lay.block_cols = (num_columns + nb - 1) / nb;
    cb->lay.num_micro = (lay.num_rows + 31) / 32;
    cudaError_t e = idx_bits == 64 ? cb_build_typed<int64_t>(cb, vals, (const int64_t *)cols, lay, stream)
                                   : cb_build_typed<int32_t>(cb, vals, (const int32_t *)cols, lay, stream);
    if (e != cudaSuccess) { cb_free(cb); return e; }
    // reserve an L2 persisting carve-out as large as one x block (best effort)
    int dev = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaGetDeviceProperties(&prop, dev) == cudaSuccess) {
        cb->num_sms = prop.multiProcessorCount;
        const size_t want = (size_t)cb->lay.block_cols * 8;
        if (prop.persistingL2CacheMaxSize > 0 && want <= (size_t)prop.accessPolicyMaxWindowSize &&
            !getenv("ELLSPMV_CUDA_NO_PERSIST")) {
            const size_t lim = want < (size_t)prop.persistingL2CacheMaxSize ? want : (size_t)prop.persistingL2CacheMaxSize;
            if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, lim) == cudaSuccess) cb->persist = true;
            cudaGetLastError();
        }
    }
    *out = cb;
    return cudaSuccess;
}

int64_t cb_bytes(const CbMatrix *cb) { return cb ? cb->bytes : 0; }
int cb_blocks(const CbMatrix *cb) { return cb ? cb->lay.num_blocks : 0; }
int64_t cb_entries(const CbMatrix *cb) { return cb ? cb->lay.total_entries : 0; }

// ---- kernel --------------------------------------------------------------------
// Persistent warps: warp w walks micro-slices w, w+W, w+2W, ... of one column
// block; lane p = the row with the p-th most entries.  The (count, lane) record
// and the start offset of the NEXT micro-slice are requested before the current
// one is processed, which removes one of the three dependent memory hops
// (record/offset -> entries -> gathers) from every task.
template <typename IdxT, bool FMA>
__global__ void __launch_bounds__(kBlockThreads)
ell_blocked_kernel(const double *__restrict__ vals, const IdxT *__restrict__ cols, const long long *__restrict__ offs,
                   const unsigned short *__restrict__ meta, const double *__restrict__ x, double *__restrict__ y,
                   int64_t num_rows, int64_t num_micro, int block, int mode /* 0: y = acc, 1: y += acc */)
{
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int64_t m = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    if (m >= num_micro) return;
    const int64_t i0 = (int64_t)block * num_micro;
    unsigned rec_next = meta[(i0 + m) * 32 + lane];
    long long base_next = offs[i0 + m];
    for (; m < num_micro; m += warps) {
        const unsigned rec = rec_next;
        long long base = base_next;
        const int64_t mn = m + warps;
        if (mn < num_micro) {                       // prefetch the next task's record and offset
            rec_next = meta[(i0 + mn) * 32 + lane];
            base_next = offs[i0 + mn];
        }
        const int cnt = rec & 0xff;
        const int64_t row = m * 32 + (rec >> 8);
        const int maxc = __shfl_sync(0xffffffffu, cnt, 0);        // lane 0 holds the longest row
        // y is only needed at the end: ask for it first so that its latency hides behind the loop
        double yold = 0.0;
        if (mode != 0 && cnt > 0) yold = y[row];
        // U diagonals per batch, all loads of a batch before its gathers.  Kept at 4 and few
        // registers on purpose: a deeper per-warp pipeline (8 + prefetch, 96 registers) measured
        // 14.3 ms on BASELINE config 4 against 9.6 ms -- with only a few hundred entries per
        // task, resident warps hide latency better (profiles/r1_c4_column_blocked.md)
        constexpr int U = 4;
        double acc = 0.0;
        for (int j = 0; j < maxc; j += U) {
            double v[U], xv[U]; int64_t c[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const bool on = cnt > j + u;
                const unsigned active = __ballot_sync(0xffffffffu, on);
                v[u] = 0.0; c[u] = -1;
                if (on) { v[u] = __ldcs(vals + base + lane); c[u] = (int64_t)__ldcs(cols + base + lane); }
                base += __popc(active);
            }
#pragma unroll
            for (int u = 0; u < U; u++) xv[u] = c[u] >= 0 ? __ldg(x + c[u]) : 0.0;
#pragma unroll
            for (int u = 0; u < U; u++)
                if (c[u] >= 0) acc = FMA ? __fma_rn(v[u], xv[u], acc) : __dadd_rn(acc, __dmul_rn(v[u], xv[u]));
        }
        if (row < num_rows) {
            if (mode == 0) y[row] = __dadd_rn(0.0, acc);
            else if (cnt > 0) y[row] = __dadd_rn(yold, acc);
        }
    }
}

template <typename IdxT, bool FMA>
static cudaError_t cb_launch_block(const CbMatrix *cb, const double *x, double *y, int64_t num_rows, int64_t num_columns,
                                   int b, int mode, unsigned grid, cudaStream_t stream)
{
    // persistent grid: exactly the CTAs that are resident at once (a multiple of the SM count)
    static int per_sm = 0;
    if (per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ell_blocked_kernel<IdxT, FMA>, kBlockThreads, 0) != cudaSuccess || n < 1) n = 8;
        per_sm = n;
    }
    const unsigned resident = (unsigned)(cb->num_sms * per_sm);
    if (grid > resident) grid = resident;
    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3(grid);
    lc.blockDim = dim3(kBlockThreads);
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    lc.attrs = attr;
    lc.numAttrs = 0;
    if (cb->persist) {
        // this block's slice of x stays in L2 while the matrix streams through
        const int64_t c0 = (int64_t)b * cb->lay.block_cols;
        int64_t c1 = c0 + cb->lay.block_cols;
        if (c1 > num_columns) c1 = num_columns;
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<double *>(x + c0);
        attr[0].val.accessPolicyWindow.num_bytes = (size_t)(c1 - c0) * 8;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        lc.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&lc, ell_blocked_kernel<IdxT, FMA>, (const double *)cb->vals, (const IdxT *)cb->cols,
                              (const long long *)cb->offs, (const unsigned short *)cb->meta, x, y, num_rows,
                              cb->lay.num_micro, b, mode);
}

cudaError_t cb_spmv(const CbMatrix *cb, bool fma, const double *x, double *y, int64_t num_rows, int64_t num_columns,
                    int beta, cudaStream_t stream)
{
    const CbLayout &L = cb->lay;
    int64_t grid = (L.num_micro * 32 + kBlockThreads - 1) / kBlockThreads;
    if (grid <= 0) return cudaSuccess;
    if (grid > 0x7fffffffLL) grid = 0x7fffffffLL;   // clipped to the resident CTA count at launch
    for (int b = 0; b < L.num_blocks; b++) {
        const int mode = (b == 0 && !beta) ? 0 : 1;
        cudaError_t e;
        if (cb->idx_bits == 64)
            e = fma ? cb_launch_block<int64_t, true>(cb, x, y, num_rows, num_columns, b, mode, (unsigned)grid, stream)
                    : cb_launch_block<int64_t, false>(cb, x, y, num_rows, num_columns, b, mode, (unsigned)grid, stream);
        else
            e = fma ? cb_launch_block<int32_t, true>(cb, x, y, num_rows, num_columns, b, mode, (unsigned)grid, stream)
                    : cb_launch_block<int32_t, false>(cb, x, y, num_rows, num_columns, b, mode, (unsigned)grid, stream);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace ellspmv
