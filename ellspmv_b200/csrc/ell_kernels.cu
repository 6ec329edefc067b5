// ell_kernels.cu -- fp64 ELLPACK y <- beta*y + A*x for sm_100a.
//
// Replaces the reference's hot loop `ellgemv` (ellspmv.c:1146-1151):
//     for i: yi = 0; for l < K: yi += a[i*K+l] * x[colidx[i*K+l]]; y[i] += yi
//
// Thread-per-row kernel (bit-exact path).  Each thread owns R consecutive
// rows of one slice of the sliced-ELL layout (common.cuh) and walks the K
// slots in the reference's order with separate __dmul_rn/__dadd_rn, so the
// rounding sequence is exactly the reference's compiled loop (mul, then add,
// left to right).  Per slot a thread issues ONE vector load
// of R values (64/128/256-bit) and ONE vector load of R indices; consecutive
// threads read consecutive addresses, so a warp's request is a contiguous
// 32*R*8-byte run.  The matrix streams bypass L1 (L1::no_allocate) and are
// marked evict-first in L2 where the ISA allows it (256-bit form), leaving
// L1/L2 to the x gather, which goes through the read-only path (ld.global.nc).
//
// This kernel is HBM-bound: 2 flops per 12..16 streamed bytes.  Tensor cores
// do not apply (gather + fp64 dot, no dense contraction).
#include "ell_thread.cuh"

namespace ellspmv {

// ---- sub-warp-per-row kernel (tolerance mode) -------------------------------
// T lanes share one row: lane j of the group takes slots j, j+T, j+2T, ...
// and the partial sums are combined with a shuffle-xor tree, so the
// summation order differs from the reference (tolerance documented in
// DESIGN.md).  Layout is the same sliced ELL with S = kBlockThreads rows.
// Within a warp, lanes [j*(32/T), (j+1)*(32/T)) hold slot-class j for 32/T
// consecutive rows, so each class reads a contiguous run.
template <typename IdxT, int T, bool FMA>
__global__ void __launch_bounds__(kBlockThreads)
ell_subwarp_kernel(const EllSpmvArgs a, int slice_rows)
{
    constexpr int RW = 32 / T;                       // rows per warp
    constexpr int WARPS = kBlockThreads / 32;
    const int K = a.rowsize;
    const int S = slice_rows;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = lane / RW, rl = lane % RW;
    // a CTA covers WARPS*RW rows per pass and S rows in total
    const int64_t slice = a.slice_begin + blockIdx.x;
    const double *vbase = a.vals + slice * S * (int64_t)K;
    const IdxT *cbase = reinterpret_cast<const IdxT *>(a.cols) + slice * S * (int64_t)K;
    const double *__restrict__ x = a.x;
    for (int r_in = warp * RW + rl; r_in < S; r_in += WARPS * RW) {
        const int64_t row = slice * S + r_in;
        double acc = 0.0;
        if (row < a.num_rows) {
#pragma unroll 4
            for (int l = j; l < K; l += T) {
                double v; int64_t c[1]; double vv[1];
                Vals<1>::ld(vbase + (int64_t)l * S + r_in, vv); v = vv[0];
                Cols<IdxT, 1>::ld(cbase + (int64_t)l * S + r_in, c);
                acc = madd<FMA>(acc, v, __ldg(x + c[0]));
            }
        }
#pragma unroll
        for (int off = RW; off < 32; off <<= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (j == 0 && row < a.num_rows) {
            if (a.ad) acc += a.ad[row] * __ldg(x + a.row_begin + row);
            double out = a.beta ? a.y[row] + acc : acc;
            a.y[row] = out;
            const int64_t g = a.row_begin + row;
            for (int p = 0; p < a.push.num_peers; p++)
                if (g >= a.push.row_lo[p] && g < a.push.row_hi[p]) a.push.x[p][g] = out;
        }
    }
}

// ---- launcher ----------------------------------------------------------------
template <typename IdxT, int R, int KU, bool FMA, int G>
static cudaError_t launch_thread_g(const EllSpmvArgs &args, bool yvec, cudaLaunchConfig_t &lc)
{
    if (args.rowlen) return cudaErrorInvalidValue;     // per-row lengths: launch_ell_spmv hands them to ell_kernels_len.cu
    if (args.sync.local_flags) {
        // the fused step hand-shake: separate instantiations, so that every other launch runs a
        // kernel without a trace of it (the whole-group PAT = 1 form serves masked handles too:
        // patinfo's low byte is the id)
        if (args.patinfo || args.patlane || args.vpat) return cudaErrorNotSupported;   // api.cu does not fuse for these handles
        if (args.patid) {
            if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, true, G, 1, false, true>, args);
            return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, false, G, 1, false, true>, args);
        }
        if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, true, G, 0, false, true>, args);
        return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, false, G, 0, false, true>, args);
    }
    if (args.patinfo) {
        if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, true, G, 2>, args);
        return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, false, G, 2>, args);
    }
    if (args.vpat) {
        // value patterns: bit-exact arithmetic only (api.cu does not look for them with FMA)
        if constexpr (!FMA) {
            if (args.patlane) {
                if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, false, true, G, 3, false, false, true>, args);
                return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, false, false, G, 3, false, false, true>, args);
            }
            if (args.patid) {
                if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, false, true, G, 1, false, false, true>, args);
                return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, false, false, G, 1, false, false, true>, args);
            }
        }
        return cudaErrorNotSupported;
    }
    if (args.patlane) {
        if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, true, G, 3>, args);
        return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, false, G, 3>, args);
    }
    if (args.patid) {
        if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, true, G, 1>, args);
        return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, false, G, 1>, args);
    }
    if (yvec) return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, true, G, 0>, args);
    return cudaLaunchKernelEx(&lc, ell_thread_kernel<IdxT, R, KU, FMA, false, G, 0>, args);
}

template <typename IdxT, int R, int KU, bool FMA>
static cudaError_t launch_thread_yvec(const EllSpmvArgs &args, int64_t num_slices, bool yvec,
                                      cudaLaunchConfig_t &lc, int gather)
{
    // only the L1-allocating read-only gather is instantiated: on B200 the other
    // flavours are never faster (tools/experiments/l2_fetch_granularity.cu,
    // profiles/r1_c4_gather.md): no_allocate loses the L1 reuse of stencil rows,
    // and for scattered rows every flavour costs one 128-byte DRAM line per gather
    (void)num_slices; (void)gather;
    return launch_thread_g<IdxT, R, KU, FMA, 0>(args, yvec, lc);
}

template <typename IdxT, int R, bool FMA>
static cudaError_t launch_thread_k(const EllSpmvArgs &args, int64_t num_slices, bool yvec,
                                   cudaLaunchConfig_t &lc, int gather)
{
    switch (args.rowsize) {
    case 5:  return launch_thread_yvec<IdxT, R, 5, FMA>(args, num_slices, yvec, lc, gather);
    case 27: return launch_thread_yvec<IdxT, R, 27, FMA>(args, num_slices, yvec, lc, gather);
    case 32: return launch_thread_yvec<IdxT, R, 32, FMA>(args, num_slices, yvec, lc, gather);
    default: return launch_thread_yvec<IdxT, R, 0, FMA>(args, num_slices, yvec, lc, gather);
    }
}

template <typename IdxT, bool FMA>
static cudaError_t launch_thread_r(int R, const EllSpmvArgs &args, int64_t num_slices, bool yvec,
                                   cudaLaunchConfig_t &lc, int gather)
{
    switch (R) {
    case 1: return launch_thread_k<IdxT, 1, FMA>(args, num_slices, yvec, lc, gather);
    case 2: return launch_thread_k<IdxT, 2, FMA>(args, num_slices, yvec, lc, gather);
    case 4: return launch_thread_k<IdxT, 4, FMA>(args, num_slices, yvec, lc, gather);
    default: return cudaErrorInvalidValue;
    }
}

template <typename IdxT, bool FMA>
static cudaError_t launch_subwarp(const EllSpmvArgs &args, int slice_rows, cudaLaunchConfig_t &lc)
{
    const int K = args.rowsize;
    // lanes per row grow with the row length: about 6..12 slots per lane
    if (K >= 192) return cudaLaunchKernelEx(&lc, ell_subwarp_kernel<IdxT, 32, FMA>, args, slice_rows);
    if (K >= 96) return cudaLaunchKernelEx(&lc, ell_subwarp_kernel<IdxT, 16, FMA>, args, slice_rows);
    if (K >= 24) return cudaLaunchKernelEx(&lc, ell_subwarp_kernel<IdxT, 8, FMA>, args, slice_rows);
    if (K >= 12) return cudaLaunchKernelEx(&lc, ell_subwarp_kernel<IdxT, 4, FMA>, args, slice_rows);
    return cudaLaunchKernelEx(&lc, ell_subwarp_kernel<IdxT, 2, FMA>, args, slice_rows);
}

static bool aligned_to(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

cudaError_t launch_ell_spmv(const EllLaunchCfg &cfg, const EllSpmvArgs &args_in,
                            int64_t num_slices, cudaStream_t stream)
{
    if (num_slices <= 0) return cudaSuccess;
    EllSpmvArgs args = args_in;
    // experiments: ELLSPMV_CUDA_PREFETCH_SLICES overrides the prefetch distance (0 = off)
    static const int prefetch_env = getenv("ELLSPMV_CUDA_PREFETCH_SLICES") ? atoi(getenv("ELLSPMV_CUDA_PREFETCH_SLICES")) : -1;
    if (prefetch_env >= 0) args.prefetch = prefetch_env;
    if (num_slices > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (cfg.kernel == kKernelLongRow) return launch_ell_longrow(cfg, args, stream);
    if ((cfg.variant & 1) && args.num_rows > 0) {
        bool handled = false;
        cudaError_t e = launch_ell_bulk(cfg, args, num_slices, stream, &handled);
        if (e != cudaSuccess || handled) return e;
    }

    cudaLaunchConfig_t lc = {};
    lc.gridDim = dim3((unsigned)num_slices);
    lc.blockDim = dim3(kBlockThreads);
    lc.dynamicSmemBytes = 0;
    lc.stream = stream;
    cudaLaunchAttribute attr[1];
    lc.attrs = attr;
    lc.numAttrs = 0;
    if (cfg.persist_x && cfg.x_bytes > 0) {
        // L2 persisting window over x: the gather target stays resident while
        // the matrix streams pass through (B200 analogue of the reference's
        // A64FX sector-cache isolation, ellspmv.c:1139-1141)
        attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
        attr[0].val.accessPolicyWindow.base_ptr = const_cast<double *>(args.x);
        attr[0].val.accessPolicyWindow.num_bytes = (size_t)cfg.x_bytes;
        attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
        attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
        lc.numAttrs = 1;
    }

    const bool i64 = cfg.idx_bits == 64;
    if (cfg.kernel == ELLSPMV_CUDA_KERNEL_WARP) {
        const int slice_rows = kBlockThreads * cfg.rows_per_thread;
        if (i64) return cfg.fma ? launch_subwarp<int64_t, true>(args, slice_rows, lc)
                                : launch_subwarp<int64_t, false>(args, slice_rows, lc);
        return cfg.fma ? launch_subwarp<int32_t, true>(args, slice_rows, lc)
                       : launch_subwarp<int32_t, false>(args, slice_rows, lc);
    }

    // vector y access needs R*8-byte alignment of y, of the row offset, and
    // of every push target
    const int R = cfg.rows_per_thread;
    bool yvec = aligned_to(args.y, 8 * (size_t)R) && (args.row_begin % R == 0);
    for (int p = 0; p < args.push.num_peers; p++) yvec = yvec && aligned_to(args.push.x[p], 8 * (size_t)R);

    if (args.rowlen) return launch_ell_thread_len(cfg, args, yvec, lc);     // the CSR view (ell_kernels_len.cu)
    const int gather = 0;
    if (i64) return cfg.fma ? launch_thread_r<int64_t, true>(R, args, num_slices, yvec, lc, gather)
                            : launch_thread_r<int64_t, false>(R, args, num_slices, yvec, lc, gather);
    return cfg.fma ? launch_thread_r<int32_t, true>(R, args, num_slices, yvec, lc, gather)
                   : launch_thread_r<int32_t, false>(R, args, num_slices, yvec, lc, gather);
}

}  // namespace ellspmv
