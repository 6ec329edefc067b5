// ell_bulk.cu -- persistent, bulk-async (TMA) staged variant of the
// thread-per-row ELL kernel (ELLSPMV_CUDA_VARIANT = 1).
//
// Same arithmetic and the same sliced layout as ell_thread_kernel
// (ell_kernels.cu); what differs is how the matrix streams reach the SM.
// A slice is one contiguous S*K*(8+idx)-byte region of HBM, so instead of
// every thread issuing its own vector loads, one thread per CTA asks the
// bulk-copy engine for the whole slice (cp.async.bulk global -> shared,
// completion counted on an mbarrier: SASS UBLKCP), NST slices ahead, while
// the CTA's threads consume the previous slice out of shared memory.  CTAs
// are persistent: the grid is a multiple of the SM count and each CTA walks
// slices b, b+G, b+2G, ...
//
// Measured against the direct-load kernel in profiles/ (r1_bulk_variant.md);
// both are HBM-bound, the default stays whichever is faster there.
#include "common.cuh"

namespace ellspmv {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a copy that never lands must not hang the GPU (results would be
// wrong and the parity tests would say so)
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (int spins = 0; !mbar_try_wait(bar, parity); spins++)
        if (spins > (1 << 24)) __trap();   // a bulk copy that never lands: fail the launch loudly, never sum unfilled shared memory
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <int R> struct SVec;
template <> struct SVec<1> {
    static __device__ __forceinline__ void ldv(const double *p, double (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void ldc(const int32_t *p, int64_t (&c)[1]) { c[0] = *p; }
    static __device__ __forceinline__ void ldc(const int64_t *p, int64_t (&c)[1]) { c[0] = *p; }
};
template <> struct SVec<2> {
    static __device__ __forceinline__ void ldv(const double *p, double (&v)[2]) {
        double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void ldc(const int32_t *p, int64_t (&c)[2]) {
        int2 t = *reinterpret_cast<const int2 *>(p); c[0] = t.x; c[1] = t.y;
    }
    static __device__ __forceinline__ void ldc(const int64_t *p, int64_t (&c)[2]) {
        longlong2 t = *reinterpret_cast<const longlong2 *>(p); c[0] = t.x; c[1] = t.y;
    }
};
template <> struct SVec<4> {
    static __device__ __forceinline__ void ldv(const double *p, double (&v)[4]) {
        double2 t0 = reinterpret_cast<const double2 *>(p)[0], t1 = reinterpret_cast<const double2 *>(p)[1];
        v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
    }
    static __device__ __forceinline__ void ldc(const int32_t *p, int64_t (&c)[4]) {
        int4 t = *reinterpret_cast<const int4 *>(p); c[0] = t.x; c[1] = t.y; c[2] = t.z; c[3] = t.w;
    }
    static __device__ __forceinline__ void ldc(const int64_t *p, int64_t (&c)[4]) {
        longlong2 t0 = reinterpret_cast<const longlong2 *>(p)[0], t1 = reinterpret_cast<const longlong2 *>(p)[1];
        c[0] = t0.x; c[1] = t0.y; c[2] = t1.x; c[3] = t1.y;
    }
};

template <int R> __device__ __forceinline__ void st_vec(double *p, const double (&v)[R]);
template <> __device__ __forceinline__ void st_vec<1>(double *p, const double (&v)[1]) { *p = v[0]; }
template <> __device__ __forceinline__ void st_vec<2>(double *p, const double (&v)[2]) {
    *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
}
template <> __device__ __forceinline__ void st_vec<4>(double *p, const double (&v)[4]) {
    reinterpret_cast<double2 *>(p)[0] = make_double2(v[0], v[1]);
    reinterpret_cast<double2 *>(p)[1] = make_double2(v[2], v[3]);
}

constexpr int kBulkHeader = 128;   // mbarriers live in front of the stage buffers

template <typename IdxT, int R, bool FMA>
__global__ void __launch_bounds__(kBlockThreads)
ell_bulk_kernel(const EllSpmvArgs a, int64_t num_slices, int nst)
{
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int S = kBlockThreads * R;
    constexpr int U = 4;
    const int K = a.rowsize;
    const uint32_t vbytes = (uint32_t)S * K * 8, cbytes = (uint32_t)S * K * sizeof(IdxT);
    const uint32_t stage_bytes = vbytes + cbytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    unsigned char *data = smem + kBulkHeader;
    const int tid = threadIdx.x;

    if (tid == 0) {
        for (int i = 0; i < nst; i++) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t stride = gridDim.x;
    const int64_t n_my = num_slices > (int64_t)blockIdx.x ? (num_slices - blockIdx.x + stride - 1) / stride : 0;
    const double *vals = a.vals;
    const IdxT *cols = reinterpret_cast<const IdxT *>(a.cols);
    const double *__restrict__ x = a.x;
    const bool yvec = (reinterpret_cast<uintptr_t>(a.y) % (8 * R)) == 0;

    auto issue = [&](int64_t i) {
        const int st = (int)(i % nst);
        const int64_t slice = a.slice_begin + blockIdx.x + i * stride;
        unsigned char *dst = data + (size_t)st * stage_bytes;
        mbar_expect_tx(&bars[st], stage_bytes);
        bulk_g2s(dst, vals + slice * S * (int64_t)K, vbytes, &bars[st]);
        bulk_g2s(dst + vbytes, cols + slice * S * (int64_t)K, cbytes, &bars[st]);
    };

    if (tid == 0)
        for (int64_t i = 0; i < nst - 1 && i < n_my; i++) issue(i);

    for (int64_t i = 0; i < n_my; i++) {
        if (tid == 0 && i + nst - 1 < n_my) {
            // the stage being refilled was read (generic proxy) in iteration i-1; the
            // __syncthreads that ended it ordered those reads, the fence hands the
            // buffer to the async proxy
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue(i + nst - 1);
        }
        const int st = (int)(i % nst);
        mbar_wait(&bars[st], (uint32_t)((i / nst) & 1));

        const int64_t slice = a.slice_begin + blockIdx.x + i * stride;
        const int64_t row0 = slice * S + (int64_t)tid * R;
        const double *sv = reinterpret_cast<const double *>(data + (size_t)st * stage_bytes) + tid * R;
        const IdxT *sc = reinterpret_cast<const IdxT *>(data + (size_t)st * stage_bytes + vbytes) + tid * R;

        if (row0 < a.num_rows) {
            const bool full = row0 + R <= a.num_rows;
            double yold[R], acc[R];
#pragma unroll
            for (int r = 0; r < R; r++) { yold[r] = 0.0; acc[r] = 0.0; }
            if (a.beta) {
                if (yvec && full) SVec<R>::ldv(a.y + row0, yold);
                else {
#pragma unroll
                    for (int r = 0; r < R; r++) if (row0 + r < a.num_rows) yold[r] = a.y[row0 + r];
                }
            }
            int l0 = 0;
#pragma unroll 1
            for (; l0 + U <= K; l0 += U) {
                double v[U][R]; int64_t c[U][R]; double xv[U][R];
#pragma unroll
                for (int u = 0; u < U; u++) {
                    SVec<R>::ldv(sv + (l0 + u) * S, v[u]);
                    SVec<R>::ldc(sc + (l0 + u) * S, c[u]);
                }
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int r = 0; r < R; r++) xv[u][r] = __ldg(x + c[u][r]);
#pragma unroll
                for (int u = 0; u < U; u++)
#pragma unroll
                    for (int r = 0; r < R; r++)
                        acc[r] = FMA ? __fma_rn(v[u][r], xv[u][r], acc[r]) : __dadd_rn(acc[r], __dmul_rn(v[u][r], xv[u][r]));
            }
#pragma unroll 1
            for (; l0 < K; l0++) {
                double v[R]; int64_t c[R];
                SVec<R>::ldv(sv + l0 * S, v);
                SVec<R>::ldc(sc + l0 * S, c);
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const double xv = __ldg(x + c[r]);
                    acc[r] = FMA ? __fma_rn(v[r], xv, acc[r]) : __dadd_rn(acc[r], __dmul_rn(v[r], xv));
                }
            }
            double out[R];
#pragma unroll
            for (int r = 0; r < R; r++) out[r] = __dadd_rn(yold[r], acc[r]);
            if (yvec && full) st_vec<R>(a.y + row0, out);
            else {
#pragma unroll
                for (int r = 0; r < R; r++) if (row0 + r < a.num_rows) a.y[row0 + r] = out[r];
            }
        }
        __syncthreads();   // everyone is done with stage st before it is refilled
    }
}

template <typename IdxT, int R, bool FMA>
static cudaError_t launch_bulk_typed(const EllLaunchCfg &cfg, const EllSpmvArgs &args, int64_t num_slices,
                                     cudaStream_t stream, bool *handled)
{
    const int S = kBlockThreads * R;
    const size_t stage_bytes = (size_t)S * args.rowsize * (8 + sizeof(IdxT));
    const size_t budget = 72 * 1024;                     // per CTA: three CTAs per SM
    int nst = (int)((budget - kBulkHeader) / stage_bytes);
    if (nst > 4) nst = 4;
    if (nst < 2) { *handled = false; return cudaSuccess; }   // slice too big to stage: caller uses direct loads
    const size_t smem = kBulkHeader + (size_t)nst * stage_bytes;
    auto kernel = ell_bulk_kernel<IdxT, R, FMA>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = (int)((227 * 1024) / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    int64_t grid = (int64_t)cfg.num_sms * per_sm;        // persistent: a multiple of the SM count
    if (grid > num_slices) grid = num_slices;
    kernel<<<(unsigned)grid, kBlockThreads, smem, stream>>>(args, num_slices, nst);
    *handled = true;
    return cudaGetLastError();
}

// returns cudaSuccess with *handled = false when this variant does not apply
cudaError_t launch_ell_bulk(const EllLaunchCfg &cfg, const EllSpmvArgs &args, int64_t num_slices,
                            cudaStream_t stream, bool *handled)
{
    *handled = false;
    if (args.push.num_peers > 0 || args.ad != nullptr || cfg.kernel != ELLSPMV_CUDA_KERNEL_THREAD) return cudaSuccess;
    const bool i64 = cfg.idx_bits == 64;
#define BULK(R_)                                                                                              \
    if (cfg.rows_per_thread == R_) {                                                                          \
        if (i64) return cfg.fma ? launch_bulk_typed<int64_t, R_, true>(cfg, args, num_slices, stream, handled) \
                                : launch_bulk_typed<int64_t, R_, false>(cfg, args, num_slices, stream, handled); \
        return cfg.fma ? launch_bulk_typed<int32_t, R_, true>(cfg, args, num_slices, stream, handled)          \
                       : launch_bulk_typed<int32_t, R_, false>(cfg, args, num_slices, stream, handled);        \
    }
    BULK(1) BULK(2) BULK(4)
#undef BULK
    return cudaSuccess;
}

}  // namespace ellspmv
