// ell_longrow.cu -- fp64 ELL y <- beta*y + A*x for FEW, LONG rows, bit-exact.
//
// The thread-per-row kernel (ell_kernels.cu) needs ~10^5 rows to keep enough loads in flight
// for HBM; a matrix with a few thousand rows of a few thousand entries each leaves most of the
// GPU idle there (profiles/r2_k_sweep.md).  The reference's loop (ellspmv.c:1146-1151) fixes the
// ORDER of the additions inside a row but says nothing about who loads the operands, so here a
// whole CTA works on a small group of rows.  Two forms share layout and arithmetic: the lock-step
// form directly below (round 2's first; kept for A/B and the parity tests,
// ELLSPMV_CUDA_LONGROW_VARIANT=0) and the ring form further down (the default: loader warps and a
// summing warp decoupled by mbarriers, 1.5-2.3x faster).  The lock-step form:
//
//   layout   row-major, exactly the reference's arrays (slice height 1): entries of a row are
//            contiguous, so 32 lanes reading 32 consecutive slots of one row is a fully
//            coalesced request at any K, and a matrix of 3 rows is not padded to 128;
//   loads    all 128 threads stream the CTA's rows tile by tile (1024 entries), gather x and
//            park the ROUNDED products a*x in shared memory;
//   sums     one lane per row then adds its row's products in slot order with __dadd_rn:
//            mul, then left-to-right adds -- the reference's rounding sequence, bit for bit;
//   overlap  two tiles are in flight in registers (values + indices of tile t+2, values + gathered
//            x of tile t+1) while tile t is being summed, one CTA barrier per tile.
//
// A row is one dependent chain of K additions whoever runs it (the CPU has the same chain), so a
// single very long row is latency-bound by construction; from a few hundred rows on the chains
// of different rows overlap and the kernel is HBM-bound like the others.  ELLSPMV_CUDA_FMA cannot
// contract here (the products are parked rounded): tolerance mode gets the same bits as the
// exact mode.
#include <stdint.h>
#include <stdlib.h>

#include "common.cuh"

namespace ellspmv {

constexpr int kLrThreads = 128;
constexpr int kLrU = 8;                          // entries per thread per tile
constexpr int kLrTile = kLrThreads * kLrU;       // 1024 entries per tile, shared by the CTA's rows

template <typename IdxT>
__device__ __forceinline__ void lr_load(const double *vp, const IdxT *cp, double &v, int64_t &c)
{
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(vp));
    if (sizeof(IdxT) == 4) { int t; asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(t) : "l"(cp)); c = t; }
    else { long long t; asm("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(t) : "l"(cp)); c = t; }
}

// rpc = rows per CTA = 1 << rshift (1..32); every row gets T_row = 1024 >> rshift slots of a tile
template <typename IdxT>
__global__ void __launch_bounds__(kLrThreads)
ell_longrow_kernel(const EllSpmvArgs a, int rshift)
{
    __shared__ double prod[2][kLrTile + 32];     // row r of a tile starts at r * (T_row + 1)
    const int K = a.rowsize;
    const int rpc = 1 << rshift;
    const int tshift = 10 - rshift;              // log2(T_row)
    const int T_row = 1 << tshift;
    const int tid = threadIdx.x;
    const int64_t row0 = (int64_t)blockIdx.x * rpc;
    const int nrows = (int)((a.num_rows - row0 < rpc) ? a.num_rows - row0 : rpc);
    const int ntiles = (K + T_row - 1) >> tshift;
    const double *__restrict__ x = a.x;
    const IdxT *cols = reinterpret_cast<const IdxT *>(a.cols);

    // this thread's entries of a tile: flat f = u*128 + tid -> row f >> tshift, slot f & (T_row-1)
    // (T_row >= 32 is a multiple of 32, so a warp's 32 lanes share one row: coalesced)
    auto issue_vc = [&](int t, double (&v)[kLrU], int64_t (&c)[kLrU]) {
#pragma unroll
        for (int u = 0; u < kLrU; u++) {
            const int f = u * kLrThreads + tid;
            const int r = f >> tshift, l = (t << tshift) + (f & (T_row - 1));
            v[u] = 0.0; c[u] = -1;
            if (t < ntiles && r < nrows && l < K) {
                const int64_t e = (row0 + r) * (int64_t)K + l;
                lr_load<IdxT>(a.vals + e, cols + e, v[u], c[u]);
            }
        }
    };
    auto issue_x = [&](const int64_t (&c)[kLrU], double (&xv)[kLrU]) {
#pragma unroll
        for (int u = 0; u < kLrU; u++) xv[u] = c[u] >= 0 ? __ldg(x + c[u]) : 0.0;
    };

    double v1[kLrU], x1[kLrU];                   // tile t+1 (at loop entry: tile 0): values and gathered x
    double v2[kLrU]; int64_t c2[kLrU];           // tile t+2 (at loop entry: tile 1): values and indices
    {
        int64_t c1[kLrU];
        issue_vc(0, v1, c1);
        issue_vc(1, v2, c2);
        issue_x(c1, x1);
    }

    // the lane that sums row r: separately stored diagonal first (ellgemvsd / ellgemv16sd orders)
    const bool summer = tid < nrows;
    const int64_t row = row0 + tid;
    // CSR view (api.cu): only the first rowlen[row] slots of a row enter the arithmetic
    const int len = (summer && a.rowlen) ? a.rowlen[row] : K;
    double acc = 0.0, dx = 0.0;
    if (summer && a.ad) {
        dx = __dmul_rn(a.ad[row], __ldg(x + a.row_begin + row));
        if (a.sd_order) acc = dx;
    }

    for (int t = 0; t < ntiles; t++) {
        double *p = prod[t & 1];
#pragma unroll
        for (int u = 0; u < kLrU; u++) {
            const int f = u * kLrThreads + tid;
            p[(f >> tshift) * (T_row + 1) + (f & (T_row - 1))] = __dmul_rn(v1[u], x1[u]);
        }
        __syncthreads();                         // tile t is parked (and tile t-1's sum is over: its buffer is free)
        // next tiles: gathers of t+1 (indices arrived during the previous sum), loads of t+2
#pragma unroll
        for (int u = 0; u < kLrU; u++) v1[u] = v2[u];
        issue_x(c2, x1);
        issue_vc(t + 2, v2, c2);
        if (summer) {
            const int n = (len - (t << tshift) < T_row) ? len - (t << tshift) : T_row;   // may be <= 0: nothing left
            // the next 8 operands leave shared memory while the current 8 are being added: the chain
            // runs at the DADD latency, not DADD + LDS
            const double *q = p + tid * (T_row + 1);
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
        }
    }
    if (!summer) return;
    if (a.ad && !a.sd_order) acc = __dadd_rn(dx, acc);
    const double yold = a.beta ? a.y[row] : 0.0;
    const double out = __dadd_rn(yold, acc);
    a.y[row] = out;
    const int64_t g = a.row_begin + row;
    for (int q = 0; q < a.push.num_peers; q++)
        if (g >= a.push.row_lo[q] && g < a.push.row_hi[q]) a.push.x[q][g] = out;
}

// ---- the ring form (default): loader warps and a summing warp decoupled by mbarriers ----------
//
// The kernel above runs in lock-step: one CTA barrier per tile, every warp waits for the indices it
// asked for one tile ago and then for warp 0's 128-step chain (ncu, 32 768 x 4096: long scoreboard
// and barrier stalls in equal parts, 3 CTAs of 160 registers per SM, 1.7 TB/s).  Here the two jobs
// never wait for each other except through a ring of parked products:
//
//   8 loader warps   stream the CTA's rows stage by stage (1024 entries, 4 per lane).  Values and
//                    indices go straight into shared memory with cp.async (LDGSTS: no registers),
//                    4 stages ahead, every lane reading back only what it asked for itself (no
//                    barrier, just cp.async.wait_group); the gathers x[c] are issued 2 stages ahead
//                    into registers (L2-only loads); the ROUNDED product a*x of stage s is parked in a
//                    ring of 2 product stages, one mbarrier arrive per warp and stage.
//   1 summing warp   lane r owns row r of the CTA (up to 32 rows): waits for a product stage, adds
//                    its row's products in slot order with __dadd_rn (the reference's rounding
//                    sequence, bit for bit -- same chain as above), hands the stage back.
//
// Same layout (row-major), same arithmetic, same y/push/diagonal/row-length handling as above.
constexpr int kLr2Loaders = 8;                           // loader warps
constexpr int kLr2Threads = (kLr2Loaders + 1) * 32;      // + the summing warp
constexpr int kLr2E = 4;                                 // entries per loader lane and stage
constexpr int kLr2Stage = kLr2Loaders * 32 * kLr2E;      // 1024 entries per stage
// Depths, measured (profiles/r2_longrow_ring.md): cp.async 4 stages ahead (3: twice as slow -- a stage then
// has one iteration to land), 2 product stages and L2-only gathers (ld.global.cg).  Shared memory and
// L1 share 256 KB per SM and every gather in flight holds an L1 line: a smaller ring and gathers that
// bypass L1 beat a deeper ring (6 stages ahead: twice as slow again; L1::no_allocate gathers: 30 % slower).
constexpr int kLr2Ahead = 4;                             // cp.async distance (stages)
constexpr int kLr2Products = 2;                          // product stages
constexpr int kLr2Gather = 2;                            // 0: ld.global.nc (__ldg), 1: nc.L1::no_allocate, 2: ld.global.cg
constexpr int kLr2ProdStride = kLr2Stage + 32;           // a row of a stage starts at r * (T_row + 1)

__device__ __forceinline__ unsigned lr_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
// a wait that never ends (a lost arrive) traps the launch after ~10 s instead of hanging the device
__device__ __forceinline__ void lr_mbar_wait(unsigned bar, unsigned parity)
{
    long long t0 = 0;
    for (int spin = 0;; spin++) {
        unsigned ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (spin == 1024) t0 = clock64();
        if (spin > 1024 && (spin & 1023) == 0 && clock64() - t0 > 20000000000LL) __trap();
    }
}
__device__ __forceinline__ void lr_mbar_arrive(unsigned bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// A per-thread constant the compiler must keep in its register: without this it re-derives staging
// and product offsets from %tid at every use (a dozen integer instructions per entry in the loop).
__device__ __forceinline__ int lr_keep(int v)
{
    asm volatile("mov.b32 %0, %0;" : "+r"(v));
    return v;
}

template <int GM>
__device__ __forceinline__ double lr_gather(const double *p)
{
    double v;
    if (GM == 1) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (GM == 2) asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else v = __ldg(p);
    return v;
}

template <typename IdxT, int kLr2Ahead, int kLr2NSP, int GM = 0, bool VEC = false>
__global__ void __launch_bounds__(kLr2Threads)
ell_longrow_ring_kernel(const EllSpmvArgs a, int rshift)
{
    constexpr int kLr2NSV = kLr2Ahead + 1;                                    // value/index staging slots
    extern __shared__ __align__(16) unsigned char lr_smem[];
    double *sv = reinterpret_cast<double *>(lr_smem);                         // [NSV][1024] values
    double *sp = sv + kLr2NSV * kLr2Stage;                                    // [NSP][1024 + 32] products
    IdxT *sc = reinterpret_cast<IdxT *>(sp + kLr2NSP * kLr2ProdStride);       // [NSV][1024] indices
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(sc + kLr2NSV * kLr2Stage);   // full[NSP], empty[NSP]

    const int K = a.rowsize;
    const int rpc = 1 << rshift;
    const int tshift = 10 - rshift;                                           // log2(T_row)
    const int T_row = 1 << tshift;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * rpc;
    const int nrows = (int)((a.num_rows - row0 < rpc) ? a.num_rows - row0 : rpc);
    const int ntiles = (K + T_row - 1) >> tshift;
    const unsigned bar0 = lr_smem_u32(bars);

    if (tid == 0) {
        for (int i = 0; i < kLr2NSP; i++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * i), "r"(kLr2Loaders));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * (kLr2NSP + i)), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp < kLr2Loaders) {
        // ---- loader: entry j of this lane is flat f = j*256 + tid of a stage -> row f >> tshift,
        // slot (stage << tshift) + (f & (T_row-1)); T_row is a multiple of 32, so the 32 lanes of an
        // instruction read 32 consecutive slots of one row.  Stages below nfull hold a real entry for
        // every lane (all rows of the CTA exist, the stage ends inside the rows): no per-entry checks
        // there; stage numbers and ring slots are running counters (no division in the loop).
        const double *__restrict__ x = a.x;
        const IdxT *cols = reinterpret_cast<const IdxT *>(a.cols);
        constexpr int kLanes = kLr2Loaders * 32;
        const int nfull = (nrows == rpc) ? (K >> tshift) : 0;
        const double *vptr[kLr2E];
        const IdxT *cptr[kLr2E];
        int lo[kLr2E], poff[kLr2E];
#pragma unroll
        for (int j = 0; j < kLr2E; j++) {
            const int f = j * kLanes + tid;
            const int r = f >> tshift, s0 = f & (T_row - 1);
            lo[j] = (r < nrows) ? s0 : K;                                     // K: never a real entry
            const int64_t e0 = (row0 + r) * (int64_t)K + s0;
            vptr[j] = a.vals + e0;
            cptr[j] = cols + e0;
            poff[j] = lr_keep(r * (T_row + 1) + s0);
        }
        // VEC (every 32-slot piece of every row starts 16-byte aligned: K a multiple of 4, or of 2 with
        // 64-bit indices): a full stage goes up in 16-byte copies that bypass L1 (cp.async.cg) -- the
        // warp's four 32-slot pieces are 64 value chunks (two per lane) and 32 / 64 index chunks (one /
        // two per lane): 3-4 LDGSTS per lane and stage instead of 8.  Lanes then read what OTHER lanes
        // of their warp asked for, hence the __syncwarp after the wait below.
        constexpr int kVC = 2, kCC = sizeof(IdxT) == 4 ? 1 : 2;               // 16-byte chunks per lane: values, indices
        const double *vsrc[kVC] = {};
        const IdxT *csrc[kCC] = {};
        int vdst[kVC] = {}, cdst[kCC] = {};
        if (VEC) {
            auto piece = [&](int seg, int64_t &e0, int &d0) {                 // piece `seg` of this warp: first entry, staging offset
                const int f = seg * kLanes + warp * 32;
                e0 = (row0 + (f >> tshift)) * (int64_t)K + (f & (T_row - 1));
                d0 = f;
            };
#pragma unroll
            for (int h = 0; h < kVC; h++) {
                const int c = lane + 32 * h;                                  // chunk c: piece c >> 4, two values at 2 * (c & 15)
                int64_t e0; int d0;
                piece(c >> 4, e0, d0);
                vsrc[h] = a.vals + e0 + 2 * (c & 15);
                vdst[h] = lr_keep(d0 + 2 * (c & 15));
            }
#pragma unroll
            for (int h = 0; h < kCC; h++) {
                const int c = lane + 32 * h;
                constexpr int per = sizeof(IdxT) == 4 ? 8 : 16, n = 16 / (int)sizeof(IdxT);   // chunks per piece, indices per chunk
                int64_t e0; int d0;
                piece(c / per, e0, d0);
                csrc[h] = cols + e0 + n * (c % per);
                cdst[h] = lr_keep(d0 + n * (c % per));
            }
        }
        int ti = 0, qi = 0;                                                   // next stage to ask for, its staging slot
        auto stage_in = [&]() {                                               // cp.async of stage ti (one group, maybe empty)
            if (VEC && ti < nfull) {
                const int64_t off = (int64_t)ti << tshift;
#pragma unroll
                for (int h = 0; h < kVC; h++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lr_smem_u32(sv + qi * kLr2Stage + vdst[h])), "l"(vsrc[h] + off) : "memory");
#pragma unroll
                for (int h = 0; h < kCC; h++)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lr_smem_u32(sc + qi * kLr2Stage + cdst[h])), "l"(csrc[h] + off) : "memory");
            } else if (ti < ntiles) {
                const int64_t off = (int64_t)ti << tshift;
                double *dv = sv + qi * kLr2Stage + tid;
                IdxT *dc = sc + qi * kLr2Stage + tid;
                const bool full = ti < nfull;
#pragma unroll
                for (int j = 0; j < kLr2E; j++) {
                    if (full || (int64_t)lo[j] + off < K) {
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(lr_smem_u32(dv + j * kLanes)), "l"(vptr[j] + off) : "memory");
                        if (sizeof(IdxT) == 4)
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(lr_smem_u32(dc + j * kLanes)), "l"(cptr[j] + off) : "memory");
                        else
                            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(lr_smem_u32(dc + j * kLanes)), "l"(cptr[j] + off) : "memory");
                    }
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            ti++;
            qi = (qi + 1 == kLr2NSV) ? 0 : qi + 1;
        };
        int tg = 0, qg = 0;                                                   // next stage to gather for, its staging slot
        auto gather = [&](double (&xv)[kLr2E]) {                              // x[c] of stage tg (its indices have landed)
            const IdxT *c = sc + qg * kLr2Stage + tid;
            if (tg < nfull) {
#pragma unroll
                for (int j = 0; j < kLr2E; j++) xv[j] = lr_gather<GM>(x + (int64_t)c[j * kLanes]);
            } else {
                const int64_t off = (int64_t)tg << tshift;
#pragma unroll
                for (int j = 0; j < kLr2E; j++) {
                    xv[j] = 0.0;
                    if (tg < ntiles && (int64_t)lo[j] + off < K) xv[j] = lr_gather<GM>(x + (int64_t)c[j * kLanes]);
                }
            }
            tg++;
            qg = (qg + 1 == kLr2NSV) ? 0 : qg + 1;
        };
        int tp = 0, qp = 0, pq = 0, pround = 0;                               // next stage to park, its staging / product slots
        auto park = [&](const double (&xv)[kLr2E]) {                          // products of stage tp -> ring, one arrive per warp
            if (pround > 0) lr_mbar_wait(bar0 + 8 * (kLr2NSP + pq), (unsigned)((pround - 1) & 1));
            double *p = sp + pq * kLr2ProdStride;
            const double *v = sv + qp * kLr2Stage + tid;
            if (tp < nfull) {
#pragma unroll
                for (int j = 0; j < kLr2E; j++) p[poff[j]] = __dmul_rn(v[j * kLanes], xv[j]);
            } else {
                const int64_t off = (int64_t)tp << tshift;
#pragma unroll
                for (int j = 0; j < kLr2E; j++)
                    if ((int64_t)lo[j] + off < K) p[poff[j]] = __dmul_rn(v[j * kLanes], xv[j]);
            }
            __syncwarp();
            if (lane == 0) lr_mbar_arrive(bar0 + 8 * pq);
            tp++;
            qp = (qp + 1 == kLr2NSV) ? 0 : qp + 1;
            if (++pq == kLr2NSP) { pq = 0; pround++; }
        };

        double xa[kLr2E], xb[kLr2E];
#pragma unroll
        for (int t = 0; t < kLr2Ahead; t++) stage_in();
        asm volatile("cp.async.wait_group %0;" ::"n"(kLr2Ahead - 2) : "memory");    // stages 0 and 1 have landed
        if (VEC) __syncwarp();
        gather(xa);
        gather(xb);
        // iteration t: ask for stage t+4, wait until stage t+2 has landed, park stage t (its gathers
        // were issued two iterations ago), issue the gathers of stage t+2 into the registers just freed
        for (int t = 0; t < ntiles; t += 2) {
            stage_in();
            asm volatile("cp.async.wait_group %0;" ::"n"(kLr2Ahead - 2) : "memory");
            if (VEC) __syncwarp();
            park(xa);
            gather(xa);
            if (t + 1 < ntiles) {
                stage_in();
                asm volatile("cp.async.wait_group %0;" ::"n"(kLr2Ahead - 2) : "memory");
                if (VEC) __syncwarp();
                park(xb);
                gather(xb);
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        return;
    }

    // ---- the summing warp: lane r adds row r's products in slot order
    const bool summer = lane < nrows;
    const int64_t row = row0 + lane;
    const int len = (summer && a.rowlen) ? a.rowlen[row] : K;                 // CSR view: only the first rowlen[row] slots count
    double acc = 0.0, dx = 0.0;
    if (summer && a.ad) {
        dx = __dmul_rn(a.ad[row], __ldg(a.x + a.row_begin + row));
        if (a.sd_order) acc = dx;
    }
    int pq = 0, pround = 0, left = len;                                       // left: slots of this row not yet added
    for (int t = 0; t < ntiles; t++) {
        lr_mbar_wait(bar0 + 8 * pq, (unsigned)(pround & 1));
        if (summer) {
            const int n = left < T_row ? left : T_row;                        // may be <= 0: nothing left
            const double *q = sp + pq * kLr2ProdStride + lane * (T_row + 1);
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
            left -= T_row;
        }
        __syncwarp();
        if (lane == 0) lr_mbar_arrive(bar0 + 8 * (kLr2NSP + pq));
        if (++pq == kLr2NSP) { pq = 0; pround++; }
    }
    if (!summer) return;
    if (a.ad && !a.sd_order) acc = __dadd_rn(dx, acc);
    const double yold = a.beta ? a.y[row] : 0.0;
    const double out = __dadd_rn(yold, acc);
    a.y[row] = out;
    const int64_t g = a.row_begin + row;
    for (int q = 0; q < a.push.num_peers; q++)
        if (g >= a.push.row_lo[q] && g < a.push.row_hi[q]) a.push.x[q][g] = out;
}

template <typename IdxT, int AHEAD, int NSP>
constexpr size_t lr2_smem_bytes()
{
    return (size_t)(AHEAD + 1) * kLr2Stage * (8 + sizeof(IdxT)) + (size_t)NSP * kLr2ProdStride * 8 + 2 * NSP * 8;
}

// Rows per CTA (1 << rshift).  A tile is 1024 entries; with r rows per CTA each row contributes
// T_row = 1024 / r consecutive slots per tile.  Two opposite costs (profiles/r2_k_sweep.md):
//   * the summing lane of a row runs T_row dependent additions per tile while the loads of the
//     next tiles are in flight: a long T_row (few rows per CTA) makes the kernel chain-bound;
//   * a row's share of a tile is one contiguous piece of HBM: a short T_row (many rows per CTA)
//     means 256-byte pieces scattered over rows that lie K*8 bytes apart, which DRAM serves badly.
// T_row = 128 (8 rows per CTA: 1 KB value pieces, 128-step chains) balances the two; short rows
// (K < 128) take more rows per CTA so that a tile is not mostly empty, and matrices with few
// rows take fewer so that there are ~4 CTAs per SM.
// the lock-step form: rows per CTA for ~4 CTAs per SM
int longrow_rshift(int64_t num_rows, int rowsize, int num_sms)
{
    int rshift = 3;
    while (rshift < 5 && (1024 >> rshift) >= 2 * rowsize) rshift++;       // K <= 64: 16 rows, K <= 32: 32 rows
    while (rshift > 0 && (num_rows >> rshift) < (int64_t)num_sms * 4) rshift--;
    return rshift;
}

// the ring form: the summing warp has 32 lanes, so up to 32 rows per CTA (32-slot pieces per row and
// stage).  What bounds a CTA is either its loaders (a stage per ~0.5 us whatever the rows per CTA) or
// the chain of its rows (1024 / rows-per-CTA dependent additions per stage): as many rows per CTA as
// still leave ~1.5 CTAs per SM (two are resident) keeps the most chains running at once
// (profiles/r2_longrow_ring.md).
int longrow_ring_rshift(int64_t num_rows, int rowsize, int num_sms)
{
    int rshift = 5;
    while (rshift > 0 && (num_rows >> rshift) < ((int64_t)num_sms * 3) / 2) rshift--;
    return rshift;
}

template <typename IdxT, int AHEAD, int NSP, int GM = 0, bool VEC = false>
static cudaError_t launch_ring(const EllSpmvArgs &args, int rshift, unsigned grid, cudaStream_t stream)
{
    static bool attr_set[64] = {};
    int dev = 0;
    cudaError_t ce = cudaGetDevice(&dev);
    if (ce != cudaSuccess) return ce;
    constexpr size_t smem = lr2_smem_bytes<IdxT, AHEAD, NSP>();
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        ce = cudaFuncSetAttribute(ell_longrow_ring_kernel<IdxT, AHEAD, NSP, GM, VEC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (ce != cudaSuccess) return ce;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    ell_longrow_ring_kernel<IdxT, AHEAD, NSP, GM, VEC><<<grid, kLr2Threads, smem, stream>>>(args, rshift);
    return cudaGetLastError();
}

template <typename IdxT>
static cudaError_t preload_ring()
{
    // an upload-time "launch" of zero rows (api.cu::warm_kernels): load the kernel's code and set its
    // shared-memory limit now, so that the first real launch is a steady-state one
    cudaFuncAttributes fa;
    cudaError_t ce = cudaFuncGetAttributes(&fa, ell_longrow_ring_kernel<IdxT, kLr2Ahead, kLr2Products, kLr2Gather, true>);
    return ce != cudaSuccess ? ce : cudaFuncGetAttributes(&fa, ell_longrow_ring_kernel<IdxT, kLr2Ahead, kLr2Products, kLr2Gather, false>);
}

cudaError_t launch_ell_longrow(const EllLaunchCfg &cfg, const EllSpmvArgs &args, cudaStream_t stream)
{
    if (args.num_rows <= 0) return cfg.idx_bits == 64 ? preload_ring<int64_t>() : preload_ring<int32_t>();
    const int sms = cfg.num_sms > 0 ? cfg.num_sms : 148;
    // read per launch (a launch here moves megabytes): the parity tests walk every form and rows-per-CTA choice
    const char *ve = getenv("ELLSPMV_CUDA_LONGROW_VARIANT"), *re = getenv("ELLSPMV_CUDA_LONGROW_RSHIFT");
    const int variant_env = ve ? atoi(ve) : 1, rshift_env = re ? atoi(re) : -1;
    const bool ring = variant_env != 0;                              // 0: the lock-step form (A/B, parity tests)
    int rshift = ring ? longrow_ring_rshift(args.num_rows, args.rowsize, sms) : longrow_rshift(args.num_rows, args.rowsize, sms);
    if (rshift_env >= 0 && rshift_env <= 5) rshift = rshift_env;     // experiments
    const int64_t grid = (args.num_rows + (1 << rshift) - 1) >> rshift;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (ring) {
        if (variant_env == 2)                                        // the first ring form: __ldg gathers, 3 product stages (A/B)
            return cfg.idx_bits == 64 ? launch_ring<int64_t, kLr2Ahead, 3, 0>(args, rshift, (unsigned)grid, stream)
                                      : launch_ring<int32_t, kLr2Ahead, 3, 0>(args, rshift, (unsigned)grid, stream);
        // 16-byte copies where every 32-slot piece of every row starts 16-byte aligned (variant 3: never)
        const int idx_per16 = cfg.idx_bits == 64 ? 2 : 4;
        const bool vec = variant_env != 3 && args.rowsize % idx_per16 == 0 &&
                         ((reinterpret_cast<uintptr_t>(args.vals) | reinterpret_cast<uintptr_t>(args.cols)) & 15) == 0;
        if (vec)
            return cfg.idx_bits == 64 ? launch_ring<int64_t, kLr2Ahead, kLr2Products, kLr2Gather, true>(args, rshift, (unsigned)grid, stream)
                                      : launch_ring<int32_t, kLr2Ahead, kLr2Products, kLr2Gather, true>(args, rshift, (unsigned)grid, stream);
        return cfg.idx_bits == 64 ? launch_ring<int64_t, kLr2Ahead, kLr2Products, kLr2Gather>(args, rshift, (unsigned)grid, stream)
                                  : launch_ring<int32_t, kLr2Ahead, kLr2Products, kLr2Gather>(args, rshift, (unsigned)grid, stream);
    }
    if (cfg.idx_bits == 64)
        ell_longrow_kernel<int64_t><<<(unsigned)grid, kLrThreads, 0, stream>>>(args, rshift);
    else
        ell_longrow_kernel<int32_t><<<(unsigned)grid, kLrThreads, 0, stream>>>(args, rshift);
    return cudaGetLastError();
}

}  // namespace ellspmv
