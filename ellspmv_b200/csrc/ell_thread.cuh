// ell_thread.cuh -- the thread-per-row ELL kernel template and its load helpers, shared by
// ell_kernels.cu (the ELL instantiations) and ell_kernels_len.cu (the instantiations with
// per-row lengths: the sliced-ELL view of a CSR matrix).  Two translation units so that nvcc
// compiles the two families in parallel.
#pragma once
#include "common.cuh"

namespace ellspmv {

// ---- vector loads of the matrix streams ----------------------------------
template <int R> struct Vals;
template <> struct Vals<1> {
    static __device__ __forceinline__ void ld(const double *p, double (&v)[1]) {
        asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v[0]) : "l"(p));
    }
};
template <> struct Vals<2> {
    static __device__ __forceinline__ void ld(const double *p, double (&v)[2]) {
        asm("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];"
                     : "=d"(v[0]), "=d"(v[1]) : "l"(p));
    }
};
template <> struct Vals<4> {
    // 256-bit load: new with sm_100 (SASS LDG.E.NA.EFL2.256.CONSTANT)
    static __device__ __forceinline__ void ld(const double *p, double (&v)[4]) {
        unsigned long long b0, b1, b2, b3;
        asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(b0), "=l"(b1), "=l"(b2), "=l"(b3) : "l"(p));
        v[0] = __longlong_as_double(b0); v[1] = __longlong_as_double(b1);
        v[2] = __longlong_as_double(b2); v[3] = __longlong_as_double(b3);
    }
};

template <typename IdxT, int R> struct Cols;
template <> struct Cols<int32_t, 1> {
    static __device__ __forceinline__ void ld(const int32_t *p, int64_t (&c)[1]) {
        int v; asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
        c[0] = v;
    }
};
template <> struct Cols<int32_t, 2> {
    static __device__ __forceinline__ void ld(const int32_t *p, int64_t (&c)[2]) {
        int v0, v1;
        asm("ld.global.nc.L1::no_allocate.v2.s32 {%0,%1}, [%2];" : "=r"(v0), "=r"(v1) : "l"(p));
        c[0] = v0; c[1] = v1;
    }
};
template <> struct Cols<int32_t, 4> {
    static __device__ __forceinline__ void ld(const int32_t *p, int64_t (&c)[4]) {
        int v0, v1, v2, v3;
        asm("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
                     : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "l"(p));
        c[0] = v0; c[1] = v1; c[2] = v2; c[3] = v3;
    }
};
template <> struct Cols<int64_t, 1> {
    static __device__ __forceinline__ void ld(const int64_t *p, int64_t (&c)[1]) {
        long long v; asm("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
        c[0] = v;
    }
};
template <> struct Cols<int64_t, 2> {
    static __device__ __forceinline__ void ld(const int64_t *p, int64_t (&c)[2]) {
        long long v0, v1;
        asm("ld.global.nc.L1::no_allocate.v2.s64 {%0,%1}, [%2];" : "=l"(v0), "=l"(v1) : "l"(p));
        c[0] = v0; c[1] = v1;
    }
};
template <> struct Cols<int64_t, 4> {
    static __device__ __forceinline__ void ld(const int64_t *p, int64_t (&c)[4]) {
        long long v0, v1, v2, v3;
        asm("ld.global.nc.L1::no_allocate.L2::evict_first.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(v0), "=l"(v1), "=l"(v2), "=l"(v3) : "l"(p));
        c[0] = v0; c[1] = v1; c[2] = v2; c[3] = v3;
    }
};

// ---- y vector access -------------------------------------------------------
template <int R> struct YVec;
template <> struct YVec<1> {
    static __device__ __forceinline__ void ld(const double *p, double (&v)[1]) { v[0] = *p; }
    static __device__ __forceinline__ void st(double *p, const double (&v)[1]) { *p = v[0]; }
};
template <> struct YVec<2> {
    static __device__ __forceinline__ void ld(const double *p, double (&v)[2]) {
        double2 t = *reinterpret_cast<const double2 *>(p); v[0] = t.x; v[1] = t.y;
    }
    static __device__ __forceinline__ void st(double *p, const double (&v)[2]) {
        *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]);
    }
};
template <> struct YVec<4> {
    static __device__ __forceinline__ void ld(const double *p, double (&v)[4]) {
        double2 t0 = reinterpret_cast<const double2 *>(p)[0];
        double2 t1 = reinterpret_cast<const double2 *>(p)[1];
        v[0] = t0.x; v[1] = t0.y; v[2] = t1.x; v[3] = t1.y;
    }
    static __device__ __forceinline__ void st(double *p, const double (&v)[4]) {
        reinterpret_cast<double2 *>(p)[0] = make_double2(v[0], v[1]);
        reinterpret_cast<double2 *>(p)[1] = make_double2(v[2], v[3]);
    }
};

// ---- the x gather --------------------------------------------------------------
// G = 0: ld.global.nc, L1-allocating (stencil-like matrices: neighbouring rows
//        share x lines, L1/L2 serve most of the gather)
// G = 1: ld.global.nc.L1::no_allocate (scattered matrices: an L1 fill pulls whole
//        128-byte lines for 8 useful bytes; see profiles/r1_c4_gather.md)
// G = 2: ld.global.cg (L2 only)
template <int G> __device__ __forceinline__ double ldx(const double *p);
template <> __device__ __forceinline__ double ldx<0>(const double *p) { return __ldg(p); }
template <> __device__ __forceinline__ double ldx<1>(const double *p) {
    double v; asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}
template <> __device__ __forceinline__ double ldx<2>(const double *p) {
    double v; asm("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p)); return v;
}

template <bool FMA>
__device__ __forceinline__ double madd(double acc, double a, double x) {
    if (FMA) return __fma_rn(a, x, acc);
    return __dadd_rn(acc, __dmul_rn(a, x));   // mul, then add: the reference's rounding
}

// slots handled per software-pipelined batch: all loads of a batch are
// issued before any use, giving U*(1+1) streamed vector loads and U*R
// gathers in flight per thread
template <int R> struct Batch { static constexpr int U = (R == 4) ? 4 : (R == 2 ? 6 : 8); };

// ---- thread-per-row kernel ------------------------------------------------
// KU > 0: K known at compile time (fully unrolled); KU == 0: run-time K.
// YVEC: y (and every push target) may be accessed with R-wide vectors.
// PAT: the handle has offset patterns (pattern.cu).  A warp whose 32*R rows share one
// offset vector d[] computes col = row + d[l] from the dictionary (a uniform load that
// lives in L1) and never touches its lines of the index stream.  PAT = 1: whole groups only
// (the default); PAT = 2: groups may carry a few deviating lanes (ELLSPMV_CUDA_PATTERN_MASKS,
// opt-in: measured slower on the BASELINE shapes, profiles/r2_offset_patterns.md); PAT = 3: one
// pattern id per THREAD, so a grid-boundary row sits in the same warp as its interior
// neighbours with its own offset vector -- same instruction stream as PAT = 1, the dictionary
// load just stops being warp-uniform (2-3 distinct L1 lines in a mixed warp).
// VPAT (with PAT = 1 or 3): value patterns -- the dictionary entry carries the K coefficients
// too (a constant-coefficient stencil), so a patterned thread reads neither stream from HBM:
// what is left is the x gather and the y store.
// LEN: rows carry their own length (a.rowlen): slots past it are loaded but never enter the
// arithmetic.  This is the CSR view (csrgemv has no padded slots, csrspmv.c:1588-1593); only
// instantiated for R = 1, PAT = 0 / 1 / 3, K = 5 / 27 / run-time (ell_kernels_len.cu).
template <typename IdxT, int R, int KU, bool FMA, bool YVEC, int G, int PAT, bool LEN = false, bool SYNC = false, bool VPAT = false>
__global__ void __launch_bounds__(kBlockThreads)
ell_thread_kernel(const EllSpmvArgs a)
{
    constexpr int S = kBlockThreads * R;
    constexpr int U = Batch<R>::U;
    const int K = KU > 0 ? KU : a.rowsize;
    const int64_t slice = a.slice_begin + blockIdx.x;
    const int64_t row0 = slice * S + (int64_t)threadIdx.x * R;   // shard-local

    // L2 prefetch (only launched with it when the matrix has offset patterns): one thread asks
    // the bulk-copy engine to pull the value stream of the slice `prefetch` CTAs ahead into L2
    // (cp.async.bulk.prefetch.L2, SASS UBLKPF.L2: no registers, no shared memory).  Without the
    // index stream the kernel is latency-bound -- the register file caps the bytes its loads keep
    // in flight -- and these requests are in flight on top of them: config 3 2.55 -> 2.09 ms,
    // config 2 0.668 -> 0.609 ms, flat from 64 to 300 slices ahead, worse beyond ~2000 (L2
    // thrash) and useless for kernels that are already HBM- or gather-bound
    // (profiles/r1_offset_patterns.md).
    if (PAT && a.prefetch > 0 && threadIdx.x == 0) {
        const int64_t ps = slice + a.prefetch;
        if ((ps + 1) * S <= a.num_rows)
            asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;"
                         :: "l"(a.vals + ps * S * (int64_t)K), "r"((unsigned)(S * K * 8)) : "memory");
    }
    // Threads past the shard's last row leave first (no CTA-wide barrier follows anywhere below:
    // the step synchronisation works warp by warp).
    if (row0 >= a.num_rows) return;

    // fused step synchronisation (row-sharded y -> x loop): a warp of a CTA that reads halo columns
    // or pushes into a peer's vector first waits until those peers have finished the previous step
    // (their pushes have landed here, and they no longer read the vector this step overwrites).
    // Warps of interior CTAs never wait, so the flag round trip over NVLink hides behind them.
    // (SYNC is a template parameter: launches without the hand-shake run the kernel without any of it)
    constexpr bool synced = SYNC;
    unsigned live = 0xffffffffu;                    // the lanes of this warp that own rows (taken while converged)
    bool boundary = false;                          // CTA-uniform: this slice pushes to a peer or reads halo columns
    if (synced) {
        live = __activemask();
        if (a.sync.num_ranges >= 0) {
            for (int i = 0; i < a.sync.num_ranges; i++)
                boundary = boundary || (slice >= a.sync.range_lo[i] && slice < a.sync.range_hi[i]);
        } else {
            boundary = a.sync.table[slice] != 0;
        }
        if (boundary && !a.sync.debug_nowait) {
            if ((threadIdx.x & 31) == 0) {
                const long long want = a.sync.epoch - 1;
                const long long t0 = clock64();
                for (int p = 0; p < a.sync.num_peers; p++) {
                    const long long *src = a.sync.local_flags + a.sync.peer_rank[p];
                    long long seen;
                    for (;;) {
                        asm volatile("ld.acquire.sys.global.s64 %0, [%1];" : "=l"(seen) : "l"(src) : "memory");
                        if (seen >= want) break;
                        if (clock64() - t0 > 40000000000LL) {   // ~20 s: a peer died; do not hang the GPU
                            if (a.sync.error) *a.sync.error = 1 + a.sync.peer_rank[p];
                            break;
                        }
                        __nanosleep(a.sync.poll_ns);
                    }
                }
            }
            __syncwarp(live);                 // nobody gathers before lane 0 has seen the flags
        }
    }

    const int64_t base = slice * S * (int64_t)K + (int64_t)threadIdx.x * R;
    const double *vp = a.vals + base;
    const IdxT *cp = reinterpret_cast<const IdxT *>(a.cols) + base;
    const double *__restrict__ x = a.x;

    // warp-uniform: this warp's pattern (or none); rowg = the row's global index
    const long long *__restrict__ prow = nullptr;
    const double *__restrict__ vrow = nullptr;      // VPAT: the pattern's coefficients
    const int64_t rowg = a.row_begin + row0;
    // A patterned group may hold a few lanes whose rows deviate from its pattern (a grid
    // boundary); they are flagged in the group's mask.  The main loop below stays the plain
    // warp-uniform two-way branch: a flagged lane runs along on the columns of the group's first
    // regular lane (valid addresses, its result is thrown away); afterwards the WHOLE warp
    // recomputes each flagged row from the explicit index stream -- lane l loads slot l, the rounded
    // products are handed to everybody by shuffles and added in slot order -- one round of loads
    // per 32 slots instead of the group's 128/256-byte index line per slot.
    unsigned pmask = 0;           // warp-uniform: the lanes of this group that deviate
    int64_t rowp = rowg;          // the row the pattern's offsets are applied to
    if (PAT) {
        if (PAT == 2) {
            const int64_t grp = (slice * kBlockThreads + threadIdx.x) >> 5;
            const unsigned long long info = __ldg(a.patinfo + grp);      // id and mask in one load
            const unsigned pid = (unsigned)(info & 0xffull);
            if (pid != 0xffu) {
                prow = a.pat + (int64_t)pid * K;
                pmask = (unsigned)(info >> 32);
                if (pmask != 0u) {
                    const int64_t lead = __shfl_sync(0xffffffffu, rowg, __ffs(~pmask) - 1);
                    if ((pmask >> (threadIdx.x & 31)) & 1u) rowp = lead;
                }
            }
        } else if (PAT == 3) {
            // pattern.cu writes 0xff into all 32 ids of a group or into none: the branch in
            // load_cols stays warp-uniform without a vote
            const unsigned pid = __ldg(a.patlane + slice * kBlockThreads + threadIdx.x);
            if (pid != 0xffu) { prow = a.pat + (int64_t)pid * K; if (VPAT) vrow = a.vpat + (int64_t)pid * K; }
        } else {
            const unsigned pid = __ldg(a.patid + ((slice * kBlockThreads + threadIdx.x) >> 5));
            if (pid != 0xffu) { prow = a.pat + (int64_t)pid * K; if (VPAT) vrow = a.vpat + (int64_t)pid * K; }
        }
    }
    auto load_vals = [&](int l, double (&v)[R]) {
        if (VPAT && prow) {
            const double av = __ldg(vrow + l);
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = av;
        } else {
            Vals<R>::ld(vp + (int64_t)l * S, v);
        }
    };
    auto load_cols = [&](int l, int64_t (&c)[R]) {
        if (PAT && prow) {
            const int64_t c0 = (PAT == 2 ? rowp : rowg) + __ldg(prow + l);
#pragma unroll
            for (int r = 0; r < R; r++) c[r] = c0 + r;
        } else {
            Cols<IdxT, R>::ld(cp + (int64_t)l * S, c);
        }
    };

    const bool full = row0 + R <= a.num_rows;
    double yold[R];
#pragma unroll
    for (int r = 0; r < R; r++) yold[r] = 0.0;
    if (a.beta) {
        if (YVEC && full) YVec<R>::ld(a.y + row0, yold);
        else {
#pragma unroll
            for (int r = 0; r < R; r++) if (row0 + r < a.num_rows) yold[r] = a.y[row0 + r];
        }
    }

    // separately stored diagonal (reference ellgemvsd / ellgemv16sd,
    // ellspmv.c:1173-1178, 1201-1219): dx = ad[i]*x[i], x[i] at the row's GLOBAL index
    const double *__restrict__ ad = a.ad;
    double dx[R];
#pragma unroll
    for (int r = 0; r < R; r++) dx[r] = 0.0;
    if (ad) {
        double d[R];
#pragma unroll
        for (int r = 0; r < R; r++) d[r] = 0.0;
        if (full) YVec<R>::ld(ad + row0, d);
        else {
#pragma unroll
            for (int r = 0; r < R; r++) if (row0 + r < a.num_rows) d[r] = ad[row0 + r];
        }
#pragma unroll
        for (int r = 0; r < R; r++)
            if (row0 + r < a.num_rows) dx[r] = __dmul_rn(d[r], __ldg(x + a.row_begin + row0 + r));
    }

    double acc[R];
#pragma unroll
    for (int r = 0; r < R; r++) acc[r] = (ad && a.sd_order) ? dx[r] : 0.0;

    int len = K, kmax = K;
    if (LEN) {
        len = a.rowlen[row0];
        kmax = __reduce_max_sync(__activemask(), len);      // slots past the warp's longest row are not even loaded
    }

    if (KU > 0) {
#pragma unroll
        for (int l0 = 0; l0 < KU; l0 += U) {
            double v[U][R]; int64_t c[U][R]; double xv[U][R];
#pragma unroll
            for (int u = 0; u < U; u++) if (l0 + u < KU) {
                load_vals(l0 + u, v[u]);
                load_cols(l0 + u, c[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) if (l0 + u < KU) {
#pragma unroll
                for (int r = 0; r < R; r++) xv[u][r] = ldx<G>(x + c[u][r]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) if (l0 + u < KU) {
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (!LEN || l0 + u < len) acc[r] = madd<FMA>(acc[r], v[u][r], xv[u][r]);
            }
        }
    } else {
        int l0 = 0;
        const int Kl = LEN ? kmax : K;
#pragma unroll 1
        for (; l0 + U <= Kl; l0 += U) {
            double v[U][R]; int64_t c[U][R]; double xv[U][R];
#pragma unroll
            for (int u = 0; u < U; u++) {
                load_vals(l0 + u, v[u]);
                load_cols(l0 + u, c[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int r = 0; r < R; r++) xv[u][r] = ldx<G>(x + c[u][r]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (!LEN || l0 + u < len) acc[r] = madd<FMA>(acc[r], v[u][r], xv[u][r]);
            }
        }
        // the last K mod U slots as ONE guarded batch: their loads are in flight together (a
        // slot-by-slot tail is K mod U dependent round trips -- all of a 5-entry row's time)
        if (l0 < Kl) {
            double v[U][R]; int64_t c[U][R]; double xv[U][R];
#pragma unroll
            for (int u = 0; u < U; u++) if (l0 + u < Kl) {
                load_vals(l0 + u, v[u]);
                load_cols(l0 + u, c[u]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) if (l0 + u < Kl) {
#pragma unroll
                for (int r = 0; r < R; r++) xv[u][r] = ldx<G>(x + c[u][r]);
            }
#pragma unroll
            for (int u = 0; u < U; u++) if (l0 + u < Kl) {
#pragma unroll
                for (int r = 0; r < R; r++)
                    if (!LEN || l0 + u < len) acc[r] = madd<FMA>(acc[r], v[u][r], xv[u][r]);
            }
        }
    }

    // the flagged lanes of a patterned group: their rows again, from the explicit indices, by the
    // whole warp (warp-uniform control flow: pmask is the same in every lane)
    if (PAT == 2 && pmask != 0u) {
        const int lane = threadIdx.x & 31;
        for (unsigned rest = pmask; rest != 0u; rest &= rest - 1) {
            const int fl = __ffs(rest) - 1;
            // element (row of lane fl, slot 0) of this slice: lanes are R rows apart
            const int64_t fbase = base + (int64_t)(fl - lane) * R;
#pragma unroll
            for (int r = 0; r < R; r++) {
                double accf = __shfl_sync(0xffffffffu, (ad && a.sd_order) ? dx[r] : 0.0, fl);
#pragma unroll 1
                for (int l0 = 0; l0 < K; l0 += 32) {
                    const int l = l0 + lane;
                    double pv = 0.0, px = 0.0;
                    if (l < K) {
                        double v1[1]; int64_t c1[1];
                        Vals<1>::ld(a.vals + fbase + r + (int64_t)l * S, v1);
                        Cols<IdxT, 1>::ld(reinterpret_cast<const IdxT *>(a.cols) + fbase + r + (int64_t)l * S, c1);
                        pv = v1[0];
                        px = ldx<G>(x + c1[0]);
                        if (!FMA) pv = __dmul_rn(pv, px);          // the rounded product, as in madd
                    }
                    const int n = K - l0 < 32 ? K - l0 : 32;
                    for (int j = 0; j < n; j++) {
                        const double vj = __shfl_sync(0xffffffffu, pv, j);
                        if (FMA) accf = __fma_rn(vj, __shfl_sync(0xffffffffu, px, j), accf);
                        else accf = __dadd_rn(accf, vj);
                    }
                }
                if (lane == fl) acc[r] = accf;
            }
        }
    }

    // y[i] += yi (beta=1) or y[i] = 0 + yi (beta=0; the add keeps -0 -> +0
    // exactly like "y=0; y+=yi" on the CPU)
    double out[R];
    if (ad && !a.sd_order) {
#pragma unroll
        for (int r = 0; r < R; r++) acc[r] = __dadd_rn(dx[r], acc[r]);   // ad*x + yi
    }
#pragma unroll
    for (int r = 0; r < R; r++) out[r] = __dadd_rn(yold[r], acc[r]);

    if (YVEC && full) {
        YVec<R>::st(a.y + row0, out);
    } else {
#pragma unroll
        for (int r = 0; r < R; r++) if (row0 + r < a.num_rows) a.y[row0 + r] = out[r];
    }

    // fused exchange: store the fresh y entries straight into the peers'
    // next-x vectors (peer-mapped HBM over NVLink), restricted to the row
    // range each peer references
    const int np = a.push.num_peers;
    if (np > 0) {
        const int64_t g0 = a.row_begin + row0;
        for (int p = 0; p < np; p++) {
            double *px = a.push.x[p];
            const int64_t lo = a.push.row_lo[p], hi = a.push.row_hi[p];
            if (YVEC && full && g0 >= lo && g0 + R <= hi) {
                YVec<R>::st(px + g0, out);
            } else {
#pragma unroll
                for (int r = 0; r < R; r++) {
                    const int64_t g = g0 + r;
                    if (row0 + r < a.num_rows && g >= lo && g < hi) px[g] = out[r];
                }
            }
        }
    }

    // completion count, per warp, of the BOUNDARY slices only -- the ones that push to a peer or
    // read halo columns: when the last of their warps gets here, every push of this rank is visible
    // system-wide and nothing of this rank reads the halo any more, which is all a peer needs to
    // know (interior slices touch neither).  A counter bumped by every warp of the launch cost
    // 1.3 ms per step on the 8192^2 shard (a million same-address atomics); this one sees ~10^2.
    if (synced && boundary) {
        __syncwarp(live);                           // every lane's stores (y and the pushes) are issued
        if ((threadIdx.x & 31) == 0) {
            __threadfence_system();
            const unsigned prev = atomicAdd(a.sync.done, 1u);
            if (prev == a.sync.total_warps - 1) {
                *a.sync.done = 0;                       // every warp has counted: ready for the next launch
                __threadfence_system();
                for (int p = 0; p < a.sync.num_peers; p++)
                    asm volatile("st.release.sys.global.s64 [%0], %1;"
                                 ::"l"(a.sync.peer_flags[p] + a.sync.rank), "l"(a.sync.epoch) : "memory");
            }
        }
    }
}

}  // namespace ellspmv
