// ell_staged.cu -- column-blocked ELL with the x gather STAGED through HBM
// (ELLSPMV_CUDA_STAGED_GATHER): the bit-exact answer to matrices whose x does not
// fit in L2.
//
// Why: on B200 a random 8-byte gather that misses L2 costs a ~100-byte line fill
// (profiles/r1_c4_gather.md), so the plain kernel runs BASELINE config 4 (random
// 50M x 32, x = 400 MB) at 9x its algorithmic bytes.  ell_blocked.cu fixes the
// locality by summing block after block, which changes the association of the
// row sums (tolerance mode).  Here the two halves of `a*x[col]` are separated
// instead, so that the arithmetic keeps the reference's exact order
// (ellspmv.c:1146-1151):
//
//   phase 1 (gather + multiply): the entries are stored a second time, sorted by
//     (column block, slice): column index and value.  One launch per column
//     block walks its run -- two flat, perfectly coalesced streams -- gathers
//     x[col] -- the live part of x is then that block's slice (<= 48 MB), which
//     stays in L2 -- and writes the ROUNDED product a*x (__dmul_rn) as a flat
//     stream xg.
//   phase 2 (sum): one CTA per slice.  For a slice the products are nb
//     contiguous runs of xg (one per column block); one thread fetches them into
//     shared memory with bulk-async copies (cp.async.bulk + mbarrier, SASS
//     UBLKCP).  A 16-bit `pos` stream in the sliced-ELL layout says where entry
//     (row, slot) landed, and every thread then adds the products of ITS row,
//     slot 0..K-1, with __dadd_rn: mul, then left-to-right adds -- the same
//     roundings as ell_thread_kernel, bit for bit.
//   (Round 1 parked x[col] and multiplied in phase 2; moving the multiplication
//   takes the 8-byte value stream out of the HBM-bound phase 2 -- 18 -> 10 B per
//   entry -- into phase 1, which is bound by its gathers, not by HBM:
//   profiles/r2_staged_gather.md.  ELLSPMV_CUDA_FMA cannot contract here, the
//   products are parked rounded: tolerance mode gets the exact mode's bits.)
//
// Measured on BASELINE config 4 (profiles/r1_staged_gather.md): 13.6 ms against
// 28.2 ms for the direct gather, same bits.  Phase 2 runs at 103 % of the measured
// HBM copy peak (4.4 ms, ncu DRAM traffic = its algorithmic 29.7 GB).  Phase 1
// (9.2 ms) keeps the SM -> L2 request port 91 % busy (every gathered value is its
// own line request) at 174 G gathers/s, HBM only 28 % busy.  A pure-gather
// micro-benchmark reaches ~300 G/s on the same path, but a persistent phase 1
// with 2-4x the loads in flight measures the same 176 G/s: the index stream in
// and the value stream out share that port (profiles/r1_gather_paths.md).  Running
// phase 2 of one row panel next to phase 1 of the next (two streams, priorities)
// was tried and LOSES (14.7-17.7 ms): not kept.
// An L2 persisting window over the x block changes nothing here (13.62 ms with
// and without): not used.
//
// Bytes per stored entry: phase 1 reads idx (4/8) and writes 8; phase 2 reads
// 8 (value) + 2 (pos) + 8 (xg) -> 30 B for 32-bit indices, against ~112 B of
// DRAM traffic per entry for the direct gather.  Device memory: +(idx + 2 + 8)
// bytes per entry.
#include <cub/cub.cuh>

#include "common.cuh"

namespace ellspmv {

constexpr int kSgMaxBlocks = 64;
constexpr int kSgHeader = 16;                 // the mbarrier in front of the staging buffer
constexpr int kSgMaxSmem = 200 * 1024;

struct SgMatrix {
    int idx_bits = 32;
    int num_blocks = 0;
    int64_t block_cols = 0;
    int64_t num_slices = 0;
    int slice_rows = 0, rowsize = 0;
    int64_t total = 0;                        // entries of gcols / xg (segments padded to even length)
    void *gcols = nullptr;                    // column indices sorted by (block, slice)
    double *gvals = nullptr;                  // the values in the same order
    long long *seg = nullptr;                 // num_blocks * num_slices + 1 segment starts
    unsigned short *pos = nullptr;            // sliced-ELL layout: index into the slice's staging buffer
    double *xg = nullptr;                     // rounded products a*x[col], same order as gcols
    long long run_start[kSgMaxBlocks + 1];    // host copy: where column block b's run starts (seg[b * num_slices])
    int64_t bytes = 0;
    size_t smem = 0;                          // dynamic shared memory of phase 2
};

// Segment (slice s, block b) starts at seg[b * num_slices + s]: blocks outermost, so
// that a block's run and a segment's end (the next index) are both contiguous.
__host__ __device__ __forceinline__ int64_t sg_seg_index(int64_t s, int b, int64_t ns) { return (int64_t)b * ns + s; }

void sg_free(SgMatrix *sg)
{
    if (!sg) return;
    cudaFree(sg->gcols); cudaFree(sg->gvals); cudaFree(sg->seg); cudaFree(sg->pos); cudaFree(sg->xg);
    delete sg;
}
int64_t sg_bytes(const SgMatrix *sg) { return sg ? sg->bytes : 0; }
int sg_launches(const SgMatrix *sg) { return sg ? sg->num_blocks + 1 : 0; }

// ---- build ---------------------------------------------------------------------
template <typename IdxT>
__global__ void __launch_bounds__(kBlockThreads)
sg_count_kernel(const IdxT *__restrict__ cols, int S, int K, int nb, int64_t block_cols, int64_t ns,
                long long *__restrict__ sizes)
{
    __shared__ int cnt[kSgMaxBlocks];
    const int64_t s = blockIdx.x;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) cnt[b] = 0;
    __syncthreads();
    const IdxT *c = cols + s * S * (int64_t)K;
    for (int i = threadIdx.x; i < S * K; i += blockDim.x) atomicAdd(&cnt[(int)((int64_t)c[i] / block_cols)], 1);
    __syncthreads();
    for (int b = threadIdx.x; b < nb; b += blockDim.x) sizes[sg_seg_index(s, b, ns)] = (cnt[b] + 1) & ~1;
}

template <typename IdxT>
__global__ void __launch_bounds__(kBlockThreads)
sg_fill_kernel(const IdxT *__restrict__ cols, const double *__restrict__ vals, int S, int K, int nb, int64_t block_cols,
               int64_t ns, const long long *__restrict__ seg, IdxT *__restrict__ gcols, double *__restrict__ gvals,
               unsigned short *__restrict__ pos)
{
    __shared__ int cursor[kSgMaxBlocks];
    __shared__ int prefix[kSgMaxBlocks];
    __shared__ long long start[kSgMaxBlocks];
    const int64_t s = blockIdx.x;
    for (int b = threadIdx.x; b < nb; b += blockDim.x) {
        cursor[b] = 0;
        const int64_t i = sg_seg_index(s, b, ns);
        start[b] = seg[i];
        prefix[b] = (int)(seg[i + 1] - start[b]);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
        for (int b = 0; b < nb; b++) { const int len = prefix[b]; prefix[b] = run; run += len; }
    }
    __syncthreads();
    const int64_t base = s * S * (int64_t)K;
    for (int i = threadIdx.x; i < S * K; i += blockDim.x) {
        const IdxT c = cols[base + i];
        const int b = (int)((int64_t)c / block_cols);
        const int r = atomicAdd(&cursor[b], 1);
        gcols[start[b] + r] = c;
        gvals[start[b] + r] = vals[base + i];
        pos[base + i] = (unsigned short)(prefix[b] + r);
    }
}

template <typename IdxT>
static cudaError_t sg_build_typed(SgMatrix *sg, const IdxT *cols, const double *vals, cudaStream_t stream)
{
    const int64_t ns = sg->num_slices, n = (int64_t)sg->num_blocks * ns;
    const int S = sg->slice_rows, K = sg->rowsize, nb = sg->num_blocks;
    cudaError_t e;
    long long *sizes = nullptr;
    void *temp = nullptr;
    if ((e = cudaMalloc(&sizes, (size_t)(n + 1) * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&sg->seg, (size_t)(n + 1) * 8)) != cudaSuccess) { cudaFree(sizes); return e; }
    e = cudaMemsetAsync(sizes, 0, (size_t)(n + 1) * 8, stream);
    if (e == cudaSuccess) {
        sg_count_kernel<IdxT><<<(unsigned)ns, kBlockThreads, 0, stream>>>(cols, S, K, nb, sg->block_cols, ns, sizes);
        e = cudaGetLastError();
    }
    size_t temp_bytes = 0;
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(nullptr, temp_bytes, sizes, sg->seg, n + 1, stream);
    if (e == cudaSuccess) e = cudaMalloc(&temp, temp_bytes + 16);
    if (e == cudaSuccess) e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, sizes, sg->seg, n + 1, stream);
    for (int b = 0; b <= nb && e == cudaSuccess; b++)
        e = cudaMemcpyAsync(&sg->run_start[b], sg->seg + (int64_t)b * ns, 8, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(sizes);
    cudaFree(temp);
    if (e != cudaSuccess) return e;
    sg->total = sg->run_start[nb];
    const size_t ne = (size_t)sg->total + 8;                 // slack: phase 1 loads indices in aligned groups of 4
    const size_t nk = (size_t)ns * S * K;
    if ((e = cudaMalloc(&sg->gcols, ne * sizeof(IdxT))) != cudaSuccess) return e;
    if ((e = cudaMalloc(&sg->gvals, ne * 8)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&sg->pos, nk * 2)) != cudaSuccess) return e;
    if ((e = cudaMalloc(&sg->xg, ne * 8)) != cudaSuccess) return e;
    sg->bytes = (int64_t)(ne * (sizeof(IdxT) + 16) + nk * 2 + (size_t)(n + 1) * 8);
    // the padding entry of an odd segment gathers x[0] and multiplies it by 0; nobody reads the product
    if ((e = cudaMemsetAsync(sg->gcols, 0, ne * sizeof(IdxT), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(sg->gvals, 0, ne * 8, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(sg->xg, 0, ne * 8, stream)) != cudaSuccess) return e;
    sg_fill_kernel<IdxT><<<(unsigned)ns, kBlockThreads, 0, stream>>>(cols, vals, S, K, nb, sg->block_cols, ns, sg->seg,
                                                                      (IdxT *)sg->gcols, sg->gvals, sg->pos);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    return cudaStreamSynchronize(stream);
}

// ---- how scattered are the gathers? (pre-filter of KERNEL_AUTO's staged-gather trial) -------
// For sampled slices, the number of different 128-byte lines of x that the 32 lanes of a warp
// touch with one gather instruction (32 consecutive rows, same slot), averaged: ~2-3 for a
// stencil, 32 for a uniformly random matrix.
template <typename IdxT>
__global__ void __launch_bounds__(kBlockThreads)
sg_scatter_kernel(const IdxT *__restrict__ cols, EllLayout lay, int64_t stride, unsigned long long *out)
{
    const int64_t s = blockIdx.x * stride;
    if (s >= lay.num_slices) return;
    const int S = lay.slice_rows, K = lay.rowsize;
    const IdxT *c = cols + s * S * (int64_t)K;
    unsigned long long lines = 0, instr = 0;
    for (int i = threadIdx.x; i < S * K; i += blockDim.x) {     // S is a multiple of 32: a warp stays inside one slot
        const long long line = (long long)c[i] >> 4;
        const unsigned peers = __match_any_sync(0xffffffffu, line);
        if ((threadIdx.x & 31) == __ffs(peers) - 1) lines++;    // one lane per distinct line
        if ((threadIdx.x & 31) == 0) instr++;
    }
    for (int off = 16; off > 0; off >>= 1) {
        lines += __shfl_xor_sync(0xffffffffu, lines, off);
        instr += __shfl_xor_sync(0xffffffffu, instr, off);
    }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out, lines); atomicAdd(out + 1, instr); }
}

cudaError_t sg_scatter_estimate(int idx_bits, const void *cols, const EllLayout &lay, double *lines_per_gather,
                                cudaStream_t stream)
{
    *lines_per_gather = 0.0;
    if (lay.num_slices <= 0 || lay.rowsize <= 0) return cudaSuccess;
    const int64_t samples = lay.num_slices < 512 ? lay.num_slices : 512;
    const int64_t stride = lay.num_slices / samples;
    unsigned long long *d = nullptr, h[2] = {0, 0};
    cudaError_t e = cudaMalloc(&d, 16);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(d, 0, 16, stream);
    if (e == cudaSuccess) {
        if (idx_bits == 64) sg_scatter_kernel<int64_t><<<(unsigned)samples, kBlockThreads, 0, stream>>>((const int64_t *)cols, lay, stride, d);
        else sg_scatter_kernel<int32_t><<<(unsigned)samples, kBlockThreads, 0, stream>>>((const int32_t *)cols, lay, stride, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d, 16, cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    cudaFree(d);
    if (e == cudaSuccess && h[1] > 0) *lines_per_gather = (double)h[0] / (double)h[1];
    return e;
}

// device memory the staged copy of a matrix will take (to check against what is free)
int64_t sg_bytes_estimate(int idx_bits, const EllLayout &lay)
{
    return lay.entries() * (int64_t)(idx_bits / 8 + 18) + (64LL << 20);
}

static cudaError_t sg_prepare_kernels();   // per device: allow the large dynamic shared memory

// *out = nullptr (and success) when staging does not apply: x already fits the
// target, or a slice's gathered values do not fit shared memory / 16-bit positions
cudaError_t sg_build(SgMatrix **out, int idx_bits, const void *cols, const double *vals, const EllLayout &lay,
                     int64_t num_columns, int64_t target_x_bytes, cudaStream_t stream)
{
    *out = nullptr;
    if (lay.num_rows <= 0 || lay.rowsize <= 0 || num_columns <= 0) return cudaSuccess;
    int64_t nb = (num_columns * 8 + target_x_bytes - 1) / target_x_bytes;
    if (nb <= 1) return cudaSuccess;
    if (nb > kSgMaxBlocks) nb = kSgMaxBlocks;
    const int64_t stage = (int64_t)lay.slice_rows * lay.rowsize + nb;       // entries staged per slice
    if (stage > 65536 || kSgHeader + stage * 8 > kSgMaxSmem || lay.num_slices > 0x7fffffffLL) return cudaSuccess;
    SgMatrix *sg = new (std::nothrow) SgMatrix();
    if (!sg) return cudaErrorMemoryAllocation;
    sg->idx_bits = idx_bits;
    sg->num_blocks = (int)nb;
    sg->block_cols = (num_columns + nb - 1) / nb;
    sg->num_slices = lay.num_slices;
    sg->slice_rows = lay.slice_rows;
    sg->rowsize = lay.rowsize;
    sg->smem = (size_t)(kSgHeader + stage * 8);
    cudaError_t e = idx_bits == 64 ? sg_build_typed<int64_t>(sg, (const int64_t *)cols, vals, stream)
                                   : sg_build_typed<int32_t>(sg, (const int32_t *)cols, vals, stream);
    if (e == cudaSuccess) e = sg_prepare_kernels();
    if (e != cudaSuccess) { sg_free(sg); return e; }
    *out = sg;
    return cudaSuccess;
}

// ---- phase 1: xg[e] = x[gcols[e]] over one column block's run ----------------------
__device__ __forceinline__ void ld4(const int32_t *p, int64_t (&c)[4])
{
    const int4 t = __ldcs(reinterpret_cast<const int4 *>(p));
    c[0] = t.x; c[1] = t.y; c[2] = t.z; c[3] = t.w;
}
__device__ __forceinline__ void ld4(const int64_t *p, int64_t (&c)[4])
{
    const longlong2 t0 = __ldcs(reinterpret_cast<const longlong2 *>(p));
    const longlong2 t1 = __ldcs(reinterpret_cast<const longlong2 *>(p) + 1);
    c[0] = t0.x; c[1] = t0.y; c[2] = t1.x; c[3] = t1.y;
}

template <typename IdxT>
__global__ void __launch_bounds__(256)
sg_gather_kernel(const IdxT *__restrict__ gcols, const double *__restrict__ gvals, const double *__restrict__ x,
                 double *__restrict__ xg, int64_t lo, int64_t hi /* run [lo, hi), both even */)
{
    const int64_t e = ((lo >> 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
    if (e >= hi) return;
    int64_t c[4];
    ld4(gcols + e, c);                                   // aligned group of 4 (the arrays carry slack)
    const double2 a01 = __ldcs(reinterpret_cast<const double2 *>(gvals + e));
    const double2 a23 = __ldcs(reinterpret_cast<const double2 *>(gvals + e) + 1);
    const bool on0 = e >= lo, on1 = e + 2 < hi;          // which of the two pairs belong to this run
    double v[4] = {0.0, 0.0, 0.0, 0.0};
    if (on0) { v[0] = __ldg(x + c[0]); v[1] = __ldg(x + c[1]); }
    if (on1) { v[2] = __ldg(x + c[2]); v[3] = __ldg(x + c[3]); }
    // the reference's multiplication, rounded on its own (mul, THEN add: ellspmv.c:1150 as compiled)
    v[0] = __dmul_rn(a01.x, v[0]); v[1] = __dmul_rn(a01.y, v[1]);
    v[2] = __dmul_rn(a23.x, v[2]); v[3] = __dmul_rn(a23.y, v[3]);
    // one 256-bit store per group (SASS STG.E.EF.256): a lane writes a whole 32-byte sector.  Two
    // 16-byte halves -- what this kernel did first -- half-fill 32 sectors per instruction and cost
    // 0.6 gathers' worth of SM->L2 traffic per entry (176 -> 219 G gathers/s, profiles/r2_phase1_lab.md);
    // the halves remain for the first/last group of a run
    if (on0 && on1)
        asm volatile("st.global.cs.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(xg + e), "d"(v[0]), "d"(v[1]), "d"(v[2]), "d"(v[3]) : "memory");
    else {
        if (on0) __stcs(reinterpret_cast<double2 *>(xg + e), make_double2(v[0], v[1]));
        if (on1) __stcs(reinterpret_cast<double2 *>(xg + e) + 1, make_double2(v[2], v[3]));
    }
}

// ---- phase 2: the reference's row sums over the staged products ------------------------
__device__ __forceinline__ uint32_t sg_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void
sg_sum_kernel(const unsigned short *__restrict__ pos, const long long *__restrict__ seg,
              const double *__restrict__ xg, const double *__restrict__ x, double *__restrict__ y,
              const double *__restrict__ ad, int sd_order, int64_t num_rows, int64_t row_begin, int64_t ns,
              int nb, int K, int beta, const PushTargets push, const int *__restrict__ rowlen)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    double *stage = reinterpret_cast<double *>(smem + kSgHeader);
    const int S = blockDim.x, tid = threadIdx.x;
    const int64_t s = blockIdx.x;

    // segment starts and lengths of this slice, one per column block
    __shared__ long long s_o0[kSgMaxBlocks];
    __shared__ int s_len[kSgMaxBlocks];
    if (tid < nb) {
        const int64_t i = sg_seg_index(s, tid, ns);
        const long long o = seg[i];
        s_o0[tid] = o;
        s_len[tid] = (int)(seg[i + 1] - o);
    }
    __syncthreads();
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sg_smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // segments are padded to even length (16-byte copies) and sum to at most S*K + nb entries
        int total = 0;
        for (int b = 0; b < nb; b++) total += s_len[b];
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                     ::"r"(sg_smem_u32(bar)), "r"((uint32_t)total * 8u) : "memory");
        int run = 0;
        for (int b = 0; b < nb; b++) {
            const int len = s_len[b];
            if (len > 0)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sg_smem_u32(stage + run)), "l"(xg + s_o0[b]), "r"((uint32_t)len * 8u),
                               "r"(sg_smem_u32(bar)) : "memory");
            run += len;
        }
    }

    const int64_t row = s * S + tid;
    const bool live = row < num_rows;
    const int64_t base = s * S * (int64_t)K + tid;
    const unsigned short *pp = pos + base;
    double yold = 0.0, dx = 0.0;
    // CSR view: only the first rowlen[row] slots enter the arithmetic (see ell_thread_kernel's LEN)
    const int len = (rowlen && live) ? rowlen[row] : K;
    if (live && beta) yold = y[row];
    if (live && ad) dx = __dmul_rn(ad[row], __ldg(x + row_begin + row));

    // the positions of the first batch travel while the bulk copies land
    constexpr int U = 8;
    unsigned p[U];
#pragma unroll
    for (int u = 0; u < U; u++) p[u] = u < K ? (unsigned)__ldcs(pp + (int64_t)u * S) : 0u;
    __syncthreads();                       // the barrier is initialised before anyone polls it
    {
        uint32_t ok = 0;
        for (int spins = 0; !ok && spins < (1 << 24); spins++)
            asm volatile("{\n\t.reg .pred p;\n\t"
                         "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                         "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sg_smem_u32(bar)), "r"(0) : "memory");
        if (!ok) __trap();     // the staged values never arrived: fail the launch loudly instead of summing garbage
    }

    // the reference's additions, slot 0..K-1, over the products phase 1 rounded
    double acc = (ad && sd_order) ? dx : 0.0;
    int l0 = 0;
#pragma unroll 1
    for (; l0 + U <= K; l0 += U) {
        unsigned pn[U];
        const bool more = l0 + 2 * U <= K;
#pragma unroll
        for (int u = 0; u < U; u++) pn[u] = more ? (unsigned)__ldcs(pp + (int64_t)(l0 + U + u) * S) : 0u;
#pragma unroll
        for (int u = 0; u < U; u++)
            if (l0 + u < len) acc = __dadd_rn(acc, stage[p[u]]);
#pragma unroll
        for (int u = 0; u < U; u++) p[u] = pn[u];
    }
    if (l0 < K) {
        // K % U leftover slots; for K < U they sit in the preloaded registers
        if (l0 == 0) {
#pragma unroll
            for (int u = 0; u < U; u++)
                if (u < K && u < len) acc = __dadd_rn(acc, stage[p[u]]);
        } else {
#pragma unroll 1
            for (; l0 < K; l0++) {
                const double pr = stage[__ldcs(pp + (int64_t)l0 * S)];
                if (l0 < len) acc = __dadd_rn(acc, pr);
            }
        }
    }
    if (!live) return;
    if (ad && !sd_order) acc = __dadd_rn(dx, acc);
    const double out = __dadd_rn(yold, acc);
    y[row] = out;
    // fused exchange, as in ell_thread_kernel: the fresh entry goes straight into the
    // next-x vector of every peer that references this row
    const int64_t g = row_begin + row;
    for (int p = 0; p < push.num_peers; p++)
        if (g >= push.row_lo[p] && g < push.row_hi[p]) push.x[p][g] = out;
}

static cudaError_t sg_prepare_kernels()
{
    return cudaFuncSetAttribute(sg_sum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSgMaxSmem);
}

cudaError_t sg_spmv(const SgMatrix *sg, const double *x, double *y, const double *ad,
                    int sd_order, int64_t num_rows, int64_t row_begin, int beta, const PushTargets *push,
                    cudaStream_t stream, const int *rowlen)
{
    PushTargets pt;
    if (push) pt = *push; else pt.num_peers = 0;
    // phase 1, one launch per column block in stream order: while block b's run is
    // gathered, the live part of x is that block's slice, which stays in L2 by itself
    for (int b = 0; b < sg->num_blocks; b++) {
        const int64_t lo = sg->run_start[b], hi = sg->run_start[b + 1];
        if (hi <= lo) continue;
        const int64_t groups = (hi + 3) / 4 - lo / 4;
        const unsigned grid = (unsigned)((groups + 255) / 256);
        if (sg->idx_bits == 64)
            sg_gather_kernel<int64_t><<<grid, 256, 0, stream>>>((const int64_t *)sg->gcols, sg->gvals, x, sg->xg, lo, hi);
        else
            sg_gather_kernel<int32_t><<<grid, 256, 0, stream>>>((const int32_t *)sg->gcols, sg->gvals, x, sg->xg, lo, hi);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    // phase 2
    sg_sum_kernel<<<(unsigned)sg->num_slices, sg->slice_rows, sg->smem, stream>>>(
        sg->pos, sg->seg, sg->xg, x, y, ad, sd_order, num_rows, row_begin, sg->num_slices, sg->num_blocks,
        sg->rowsize, beta, pt, rowlen);
    return cudaGetLastError();
}

}  // namespace ellspmv
